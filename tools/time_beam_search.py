"""Times sl_ctc_beam_search_decode at the bench shape (64 x 626 frames, V = 29) for a few beam widths."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from speechless_b200 import _lib  # noqa: E402

lib = _lib.load()
B, T, V = 64, 626, 29
rng = np.random.default_rng(0)


def scenario(name):
    """flat: random logits x 3 (hardly any blank, a new symbol almost every frame — every candidate is in play);
    trained-like: what a converged acoustic model emits — ~150 characters over 626 frames, blanks in between,
    the frame's symbol at p ~ 0.9."""
    if name == "flat":
        logits = rng.normal(size=(B, T, V)).astype(np.float32) * 3
    else:
        logits = rng.normal(size=(B, T, V)).astype(np.float32)
        target = np.full((B, T), V - 1)
        for b in range(B):
            starts = np.sort(rng.choice(T - 2, size=150, replace=False))
            target[b, starts] = rng.integers(0, V - 1, size=150)
        logits[np.arange(B)[:, None], np.arange(T)[None, :], target] += 6.0
    return torch.softmax(torch.from_numpy(logits).cuda(), dim=-1).contiguous()


lengths = torch.full((B,), T - 1, dtype=torch.int32, device="cuda")
for name in ("flat", "trained-like"):
    probs = scenario(name)
    for width in (1, 16, 100, 128):
        top = 1
        out = torch.empty((B, top, T), dtype=torch.int32, device="cuda")
        out_len = torch.empty((B, top), dtype=torch.int32, device="cuda")
        out_logp = torch.empty((B, top), dtype=torch.float32, device="cuda")
        ws = torch.empty(lib.sl_ctc_beam_search_workspace_bytes(B, T, width), dtype=torch.uint8, device="cuda")

        def run():
            _lib.check(lib.sl_ctc_beam_search_decode(_lib.ptr(probs), _lib.ptr(lengths), _lib.ptr(out),
                                                     _lib.ptr(out_len), _lib.ptr(out_logp), B, T, V, V - 1, width, top,
                                                     0, 1, _lib.ptr(ws), ws.numel(), None))
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run()
        e1.record()
        torch.cuda.synchronize()
        print("%-12s beam_width %3d: %.3f ms per batch of %d x %d frames (mean decoded length %.1f)" % (
            name, width, e0.elapsed_time(e1) / 3, B, T, out_len.float().mean().item()))
