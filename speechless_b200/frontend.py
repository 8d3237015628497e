"""GPU audio front end: raw audio -> z-normalised, transposed mel power-level spectrograms.

Replaces the librosa pipeline behind the reference's
`LabeledExample.z_normalized_transposed_spectrogram()` (labeled_example.py:99-140) with
`sl_spectrogram` + `sl_z_normalize` (csrc/frontend.cu): periodic-Hann STFT (n_fft 512, hop 128,
center=True, reflect padding), |.|^2, 10 log10 floored at -150 dB, Slaney mel projection of the
dB values (128 bins), transpose, global z-normalisation.  The batched result lives in HBM in the
very layout `Wav2Letter` feeds its tower with (zero padded to the longest utterance).
"""
from typing import List, Sequence, Tuple

import numpy as np
import torch

from speechless_b200 import _lib
from speechless_b200._lib import check, ptr

N_FFT, HOP_LENGTH, N_MELS = 512, 128, 128


def slaney_mel_filterbank(sample_rate: int = 16000, n_fft: int = N_FFT, n_mels: int = N_MELS) -> np.ndarray:
    """What `librosa.filters.mel(sr, n_fft, n_mels)` returns (labeled_example.py:113-116): triangular
    filters on the Slaney mel scale (linear below 1 kHz, logarithmic above), area normalised,
    fmin 0, fmax sr/2.  Shape (n_mels, 1 + n_fft/2)."""
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0

    def to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)

    def to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fft_frequencies = np.linspace(0, sample_rate / 2, 1 + n_fft // 2)
    edges = to_hz(np.linspace(to_mel(0.0), to_mel(sample_rate / 2), n_mels + 2))
    widths = np.diff(edges)
    ramps = edges[:, None] - fft_frequencies[None, :]
    rising = -ramps[:-2] / widths[:-1, None]
    falling = ramps[2:] / widths[1:, None]
    weights = np.maximum(0, np.minimum(rising, falling))
    return weights * (2.0 / (edges[2:] - edges[:-2]))[:, None]


def frame_count(sample_count: int, hop_length: int = HOP_LENGTH) -> int:
    return 1 + sample_count // hop_length


class SpectrogramFrontEnd:
    def __init__(self, sample_rate: int = 16000, fourier_window_length: int = N_FFT, hop_length: int = HOP_LENGTH,
                 mel_frequency_count: int = N_MELS, device=None):
        if (fourier_window_length, hop_length, mel_frequency_count) != (N_FFT, HOP_LENGTH, N_MELS):
            raise NotImplementedError("the GPU front end is built for the reference defaults: "
                                      "fourier_window_length 512, hop_length 128, 128 mel frequencies")
        if not torch.cuda.is_available():
            raise RuntimeError("speechless_b200 needs a CUDA device (B200); there is no CPU fallback.")
        self.lib = _lib.load()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.sample_rate = sample_rate
        mel = slaney_mel_filterbank(sample_rate, N_FFT, N_MELS)
        with torch.cuda.device(self.device):
            self.mel_t = torch.from_numpy(np.ascontiguousarray(mel.T, dtype=np.float32)).to(self.device)  # (257, 128)

    def batch_on_device(self, raw_audios: Sequence[np.ndarray]) -> Tuple[torch.Tensor, List[int]]:
        """-> (B, T_max, 128) fp32 device tensor (zero beyond each utterance's frames), frame counts."""
        counts = [int(len(a)) for a in raw_audios]
        if min(counts) < 2:
            raise ValueError("audio too short")
        B, stride = len(raw_audios), max(counts)
        frames = [frame_count(n) for n in counts]
        t_max = max(frames)
        host = torch.zeros((B, stride), dtype=torch.float32, pin_memory=True)
        for row, audio in enumerate(raw_audios):
            host[row, :counts[row]] = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            audio = host.to(self.device, non_blocking=True)
            sample_counts = torch.tensor(counts, dtype=torch.int32, device=self.device)
            frame_counts = torch.tensor(frames, dtype=torch.int32, device=self.device)
            out = torch.empty((B, t_max, N_MELS), dtype=torch.float32, device=self.device)
            moments = torch.empty((B, 2), dtype=torch.float64, device=self.device)
            check(self.lib.sl_spectrogram(ptr(audio), ptr(sample_counts), ptr(self.mel_t), ptr(out), B, stride, t_max,
                                          N_FFT, HOP_LENGTH, N_MELS, stream))
            check(self.lib.sl_z_normalize(ptr(out), ptr(frame_counts), ptr(moments), B, t_max, N_MELS, stream))
        return out, frames

    def z_normalized_transposed_spectrograms(self, raw_audios: Sequence[np.ndarray]) -> List[np.ndarray]:
        """One (frames, 128) array per utterance — what the reference's per-example method returns."""
        out, frames = self.batch_on_device(raw_audios)
        host = out.cpu().numpy()
        return [host[row, :frames[row]].copy() for row in range(len(frames))]
