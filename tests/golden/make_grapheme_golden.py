"""Generates tests/golden/grapheme_reference.json by importing the REFERENCE's own
speechless/grapheme_enconding.py from /root/reference (numpy-only module; the only part of the
hot path that is importable offline).  Run in the build container:
    python tests/golden/make_grapheme_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference")
from speechless.grapheme_enconding import CtcGraphemeEncoding  # noqa: E402

english = list("abcdefghijklmnopqrstuvwxyz '")
german = english + list("äöüß")
rng = np.random.default_rng(2017)
cases = []
for alphabet in (english, german):
    g = CtcGraphemeEncoding(alphabet)
    labels = ["she wasn't three abcxyz", "a", "hello  world", "".join(alphabet)]
    decodes = []
    for _ in range(6):
        graphemes = [int(v) for v in rng.integers(0, g.grapheme_set_size, size=int(rng.integers(1, 40)))]
        # make runs likely
        graphemes = [v for v in graphemes for _ in range(int(rng.integers(1, 4)))]
        decodes.append([graphemes, g.decode_graphemes(graphemes), g.decode_graphemes(graphemes, merge_repeated=False)])
    scores = rng.random((3, 25, g.grapheme_set_size)).round(3)
    scores[0, 3] = scores[0, 2]  # repeated frame
    scores[1, :, g.ctc_blank] += 0.5  # blank heavy
    scores[2, 5, :] = 0.25  # exact tie: lowest index must win
    lengths = [25, 17, 9]
    cases.append({
        "alphabet": "".join(alphabet), "labels": labels,
        "encoded": [g.encode(l) for l in labels],
        "label_batch": g.encode_label_batch(labels).tolist(),
        "decodes": decodes,
        "prediction_batch": scores.tolist(), "prediction_lengths": lengths,
        "decoded_predictions": g.decode_prediction_batch(scores, lengths),
    })
out = Path(__file__).parent / "grapheme_reference.json"
out.write_text(json.dumps({"generator": "tests/golden/make_grapheme_golden.py", "cases": cases}, ensure_ascii=False))
print("wrote", out)
