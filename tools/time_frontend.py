"""Times the GPU audio front end (sl_spectrogram + sl_z_normalize) at the bench shape: 64 utterances x 10 s of
16 kHz audio -> (64, 1251, 128) z-normalised mel power levels, device-resident audio."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speechless_b200 import _lib  # noqa: E402
from speechless_b200._lib import check, ptr  # noqa: E402
from speechless_b200.frontend import SpectrogramFrontEnd, N_FFT, HOP_LENGTH, N_MELS  # noqa: E402

front = SpectrogramFrontEnd(device="cuda:0")
lib = _lib.load()
B, samples = 64, 160000
frames = 1 + samples // HOP_LENGTH
audio = torch.randn((B, samples), dtype=torch.float32, device="cuda")
sample_counts = torch.full((B,), samples, dtype=torch.int32, device="cuda")
frame_counts = torch.full((B,), frames, dtype=torch.int32, device="cuda")
out = torch.empty((B, frames, N_MELS), dtype=torch.float32, device="cuda")
moments = torch.empty((B, 2), dtype=torch.float64, device="cuda")


def run():
    check(lib.sl_spectrogram(ptr(audio), ptr(sample_counts), ptr(front.mel_t), ptr(out), B, samples, frames, N_FFT,
                             HOP_LENGTH, N_MELS, None))
    check(lib.sl_z_normalize(ptr(out), ptr(frame_counts), ptr(moments), B, frames, N_MELS, None))


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms_spec = timed(lambda: check(lib.sl_spectrogram(ptr(audio), ptr(sample_counts), ptr(front.mel_t), ptr(out), B, samples,
                                                 frames, N_FFT, HOP_LENGTH, N_MELS, None)))
ms_norm = timed(lambda: check(lib.sl_z_normalize(ptr(out), ptr(frame_counts), ptr(moments), B, frames, N_MELS, None)))
print("  sl_spectrogram %.3f ms, sl_z_normalize %.3f ms" % (ms_spec, ms_norm))
ms = timed(run)
total = B * frames
bytes_moved = audio.numel() * 4 + 3 * out.numel() * 4  # audio read, spectrogram write + read + write (z-norm)
print("front end: %.3f ms per batch of %d x %d frames = %.1f M frames/s; %.1f MB -> %.0f GB/s" % (
    ms, B, frames, total / ms / 1e3, bytes_moved / 1e6, bytes_moved / ms / 1e6))
