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
logits = torch.from_numpy(rng.normal(size=(B, T, V)).astype(np.float32) * 3).cuda()
probs = torch.softmax(logits, dim=-1).contiguous()
lengths = torch.full((B,), T - 1, dtype=torch.int32, device="cuda")
for width in (1, 16, 100, 128):
    top = 1
    out = torch.empty((B, top, T), dtype=torch.int32, device="cuda")
    out_len = torch.empty((B, top), dtype=torch.int32, device="cuda")
    out_logp = torch.empty((B, top), dtype=torch.float32, device="cuda")
    ws = torch.empty(lib.sl_ctc_beam_search_workspace_bytes(B, T, width), dtype=torch.uint8, device="cuda")

    def run():
        _lib.check(lib.sl_ctc_beam_search_decode(_lib.ptr(probs), _lib.ptr(lengths), _lib.ptr(out), _lib.ptr(out_len),
                                                 _lib.ptr(out_logp), B, T, V, V - 1, width, top, 0, 1, _lib.ptr(ws),
                                                 ws.numel(), None))
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    print("beam_width %3d: %.3f ms per batch of %d x %d frames (mean decoded length %.1f)" % (
        width, e0.elapsed_time(e1) / 3, B, T, out_len.float().mean().item()))
