"""CPU oracle for CTC prefix beam search — TEST INFRASTRUCTURE ONLY (nothing under
`speechless_b200/` imports it).

What it restates
----------------
`tf.nn.ctc_beam_search_decoder` as the reference calls it (net.py:444-451; the KenLM scorer of the
patched TensorFlow fork named in net.py:420-422 and README.md:17 is NOT in the tree and not
reproducible — this is the stock scorer, every expansion score = the incoming probability).
The algorithm lives in TensorFlow (`tensorflow/core/util/ctc/ctc_beam_search.h`, TF 1.x; unpinned
third-party dependency, absent from /root/reference), so its published behaviour is restated:

* a tree of label prefixes; every active prefix carries log P(prefix, last frame blank),
  log P(prefix, last frame its last label) and their sum;
* per frame, active prefixes are continued (own label repeated / a blank appended; the
  "last label" mass also receives the parent's mass — the parent's blank mass only, when the
  prefix ends in a doubled label), then every active prefix spawns its not-yet-active children
  (child label == own label: from the blank mass only), and the `beam_width` best survive;
* `merge_repeated` only post-processes the winning prefix (drops a label equal to the one before
  it), which is why "A A _ A A" gives [0] with merging and [0, 0] without
  (reference speechless/test/test_ctc_decoders.py:5-9,38-39).

Pinning: the two beam-search rows of the reference's own test (test_ctc_decoders.py:38-39,
beam_width=1) are checked in tests/test_beam_search.py; with a beam wide enough to hold every
prefix the search is exact and is checked against brute-force enumeration of all V^T paths.
"""
from itertools import product
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

LOG_ZERO = -np.inf


def _lse(a: float, b: float) -> float:
    if a == LOG_ZERO:
        return b
    if b == LOG_ZERO:
        return a
    m = max(a, b)
    return m + np.log(np.exp(a - m) + np.exp(b - m))


class _Entry:
    __slots__ = ("parent", "label", "children", "old", "new")

    def __init__(self, parent: Optional["_Entry"], label: int):
        self.parent = parent
        self.label = label
        self.children: Dict[int, "_Entry"] = {}
        self.old = [LOG_ZERO, LOG_ZERO, LOG_ZERO]  # total, blank, label
        self.new = [LOG_ZERO, LOG_ZERO, LOG_ZERO]

    def active(self) -> bool:
        return self.new[0] != LOG_ZERO

    def sequence(self, merge_repeated: bool) -> List[int]:
        labels, prev, c = [], -1, self
        while c.parent is not None:
            if not merge_repeated or c.label != prev:
                labels.append(c.label)
            prev = c.label
            c = c.parent
        return labels[::-1]


def log_softmax(rows: np.ndarray) -> np.ndarray:
    m = rows.max(axis=-1, keepdims=True)
    return rows - m - np.log(np.exp(rows - m).sum(axis=-1, keepdims=True))


def beam_search_decode(inputs: np.ndarray, beam_width: int = 100, top_paths: int = 1, merge_repeated: bool = True,
                       blank: Optional[int] = None, tf_deactivation: bool = True) -> List[Tuple[List[int], float]]:
    """inputs: (T, V) unnormalised log-scores of ONE utterance (the reference feeds log(p + 1e-8)); they
    are log-softmax-normalised per frame like TF's `Step`.  Returns up to `top_paths` (labels, log
    probability) pairs, best first.

    tf_deactivation: TF's `Step` has an order-dependent side effect.  A prefix X that was in the beam and is
    pushed out by a better child DURING the child loop of a frame is, when the loop later reaches X's parent,
    re-scored as if it were new, fails (its fresh score cannot beat the beam bottom that displaced it) and has
    its PREVIOUS probabilities wiped ("deactivate child") — so X's own children are never proposed in that
    frame although X was a legitimate parent.  True restates TF; False is the order-independent rule "the
    beam_width best of {continued prefixes} U {children of every prefix of the beam}", which is what the CUDA
    kernel implements (it can only ADD hypotheses TF loses; identical whenever nothing is displaced before its
    parent is visited, e.g. for beam_width = 1 and for beams wide enough to hold every prefix)."""
    T, V = inputs.shape
    blank = V - 1 if blank is None else blank
    lp = log_softmax(np.asarray(inputs, dtype=np.float64))
    root = _Entry(None, -1)
    root.new = [0.0, 0.0, LOG_ZERO]
    leaves: List[_Entry] = [root]
    for t in range(T):
        branches = sorted(leaves, key=lambda e: -e.new[0])
        in_beam_before = set(branches)
        leaves = []
        for b in branches:
            b.old = list(b.new)
        for b in branches:
            if b.parent is not None:
                if b.parent.active():
                    previous = b.parent.old[1] if b.label == b.parent.label else b.parent.old[0]
                    b.new[2] = _lse(b.new[2], previous)
                b.new[2] += lp[t, b.label]
            b.new[1] = b.old[0] + lp[t, blank]
            b.new[0] = _lse(b.new[1], b.new[2])
            leaves.append(b)

        def bottom() -> _Entry:
            return min(leaves, key=lambda e: e.new[0])

        def is_candidate(total: float) -> bool:
            return total > LOG_ZERO and (len(leaves) < beam_width or total > bottom().new[0])

        for b in branches:
            if not is_candidate(b.old[0]):
                continue
            for label in range(V):
                if label == blank:
                    continue
                c = b.children.get(label)
                if c is None:
                    c = b.children[label] = _Entry(b, label)
                if c.active() or (not tf_deactivation and c in in_beam_before):
                    continue
                previous = b.old[1] if label == b.label else b.old[0]
                c.new = [lp[t, label] + previous, LOG_ZERO, lp[t, label] + previous]
                if is_candidate(c.new[0]):
                    if len(leaves) == beam_width:
                        worst = bottom()
                        worst.new = [LOG_ZERO, LOG_ZERO, LOG_ZERO]
                        leaves.remove(worst)
                    leaves.append(c)
                else:
                    c.old = [LOG_ZERO, LOG_ZERO, LOG_ZERO]
                    c.new = [LOG_ZERO, LOG_ZERO, LOG_ZERO]
    best = sorted(leaves, key=lambda e: -e.new[0])[:top_paths]
    return [(e.sequence(merge_repeated), float(e.new[0])) for e in best]


def collapse(path: Sequence[int], blank: int) -> Tuple[int, ...]:
    """CTC collapse of a frame path: merge runs, drop blanks."""
    out, prev = [], None
    for c in path:
        if c != prev and c != blank:
            out.append(c)
        prev = c
    return tuple(out)


def brute_force_labelings(inputs: np.ndarray, blank: Optional[int] = None) -> Dict[Tuple[int, ...], float]:
    """log P(labeling) for every labeling, by enumerating all V^T frame paths (tiny T, V only)."""
    T, V = inputs.shape
    blank = V - 1 if blank is None else blank
    lp = log_softmax(np.asarray(inputs, dtype=np.float64))
    totals: Dict[Tuple[int, ...], float] = {}
    for path in product(range(V), repeat=T):
        score = float(sum(lp[t, c] for t, c in enumerate(path)))
        key = collapse(path, blank)
        totals[key] = _lse(totals.get(key, LOG_ZERO), score)
    return totals
