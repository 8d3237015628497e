"""Times sl_ctc_beam_search_decode at the bench shape (64 x 626 frames, V = 29) for a few beam widths, and
sl_ctc_beam_search_decode_lm (word n-gram model inside the search) with a synthetic 20 000-word trigram model."""
import ctypes
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


# ---- the same search with a word n-gram model inside: 20 000 random words, 100 000 bigrams, 100 000 trigrams
from speechless_b200 import english_frequent_characters  # noqa: E402
from speechless_b200.language_model import ArpaLanguageModel, DeviceLanguageModel  # noqa: E402

alphabet = list(english_frequent_characters)
letters = [c for c in alphabet if c != " "]
words = sorted({"".join(rng.choice(letters, size=rng.integers(2, 9))) for _ in range(20000)})
ngrams = {("<s>",): (-99.0, -0.4), ("</s>",): (-1.3, 0.0), ("<unk>",): (-3.0, -0.1)}
for w in words:
    ngrams[(w,)] = (float(-rng.random() * 3 - 1), float(-rng.random() * 0.5))
for n, count in ((2, 100000), (3, 100000)):
    picks = rng.integers(0, len(words), size=(count, n))
    for row in picks:
        ngrams[tuple(words[i] for i in row)] = (float(-rng.random() * 2 - 0.1), float(-rng.random() * 0.4) if n == 2 else 0.0)
lm = DeviceLanguageModel(ArpaLanguageModel(ngrams, 3), alphabet, V, torch.device("cuda"))
print("language model: %d n-grams, table of %d slots, trie of %d nodes" % (
    len(ngrams), lm.tables.ngrams.shape[0], lm.tables.trie_children.shape[0]))
for name in ("flat", "trained-like"):
    probs = scenario(name)
    for width in (16, 100):
        out = torch.empty((B, 1, T), dtype=torch.int32, device="cuda")
        out_len = torch.empty((B, 1), dtype=torch.int32, device="cuda")
        out_logp = torch.empty((B, 1), dtype=torch.float32, device="cuda")
        ws = torch.empty(lib.sl_ctc_beam_search_workspace_bytes(B, T, width), dtype=torch.uint8, device="cuda")

        def run_lm():
            _lib.check(lib.sl_ctc_beam_search_decode_lm(_lib.ptr(probs), _lib.ptr(lengths), _lib.ptr(out),
                                                        _lib.ptr(out_len), _lib.ptr(out_logp), B, T, V, V - 1, width, 1,
                                                        0, 1, ctypes.addressof(lm.struct), _lib.ptr(ws), ws.numel(), None))
        run_lm()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run_lm()
        e1.record()
        torch.cuda.synchronize()
        print("%-12s with language model, beam_width %3d: %.3f ms per batch of %d x %d frames (mean decoded length %.1f)" % (
            name, width, e0.elapsed_time(e1) / 3, B, T, out_len.float().mean().item()))
