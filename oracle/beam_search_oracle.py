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

Word language model inside the search (`WordLanguageModelScorer`, PARITY UNPINNED): the reference passes
`kenlm_directory_path, kenlm_weight=.8, word_count_weight=0, valid_word_count_weight=2.3` to a patched
TensorFlow (net.py:444-451; fork github.com/timediv/tensorflow-with-kenlm, README.md:17 — absent from the
tree, not installable here).  TF's decoder exposes exactly four scorer hooks (`InitializeState`, `ExpandState`,
`GetStateExpansionScore`, `ExpandStateEnd` / `GetStateEndExpansionScore`, ctc_beam_search.h /
ctc_beam_scorer.h) and `beam_search_decode(scorer=...)` calls them where TF does.  The scorer behind them
is restated from the published design of the KenLM beam scorers of that time (three weights with these
names; Mozilla DeepSpeech 0.1 `beam_search.h`): a vocabulary trie follows the letters of the incomplete
word, a word is scored by the n-gram model when the space label arrives, and an unfinished word carries
the most pessimistic unigram score below its trie node (or the <unk> score once it has left the
vocabulary) as a look-ahead that is taken back when the word ends.  The total a finished hypothesis gets is
    kenlm_weight * ln P_lm(words </s>) + word_count_weight * #words + valid_word_count_weight * #known words,
the formula of `speechless_b200.language_model.NBestRescorer`; units (ln) and where the bonuses enter are
this restatement's choice — the fork's source is not available to check against.

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
    __slots__ = ("parent", "label", "children", "old", "new", "state")

    def __init__(self, parent: Optional["_Entry"], label: int):
        self.parent = parent
        self.label = label
        self.children: Dict[int, "_Entry"] = {}
        self.old = [LOG_ZERO, LOG_ZERO, LOG_ZERO]  # total, blank, label
        self.new = [LOG_ZERO, LOG_ZERO, LOG_ZERO]
        self.state = None  # scorer state (TF: BeamEntry::state)

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


class StockScorer:
    """TF's BaseBeamScorer: every expansion score is the incoming probability."""

    def initialize_state(self):
        return None

    def expand_state(self, from_state, from_label: int, to_label: int):
        return None

    def expansion_score(self, state, previous: float) -> float:
        return previous

    def expand_state_end(self, state):
        return state

    def end_expansion_score(self, state) -> float:
        return 0.0


def beam_search_decode(inputs: np.ndarray, beam_width: int = 100, top_paths: int = 1, merge_repeated: bool = True,
                       blank: Optional[int] = None, tf_deactivation: bool = True,
                       scorer=None) -> List[Tuple[List[int], float]]:
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
    scorer = scorer or StockScorer()
    root = _Entry(None, -1)
    root.new = [0.0, 0.0, LOG_ZERO]
    root.state = scorer.initialize_state()
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
                    b.new[2] = _lse(b.new[2], scorer.expansion_score(b.state, previous))
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
                c.state = scorer.expand_state(b.state, b.label, label)
                previous = scorer.expansion_score(c.state, b.old[1] if label == b.label else b.old[0])
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
    # TF TopPaths: every leaf's state is closed (end of sentence) before the leaves are ranked
    finished = []
    for e in leaves:
        if e.new[0] == LOG_ZERO:
            continue
        finished.append((e.new[0] + scorer.end_expansion_score(scorer.expand_state_end(e.state)), e))
    finished.sort(key=lambda pair: -pair[0])
    return [(e.sequence(merge_repeated), float(total)) for total, e in finished[:top_paths]]


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


# ---------------------------------------------------------------------------------------------------------
# word n-gram model inside the search (see the header: parity unpinned)
# ---------------------------------------------------------------------------------------------------------
LN10 = float(np.log(10.0))


class BackOffModel:
    """ARPA back-off n-gram model: {word tuple: (log10 probability, log10 back-off)}."""

    def __init__(self, ngrams: Dict[Tuple[str, ...], Tuple[float, float]]):
        self.ngrams = dict(ngrams)
        self.order = max(len(k) for k in self.ngrams)
        self.unknown = self.ngrams.get(("<unk>",), (-100.0, 0.0))[0]

    def knows(self, word: str) -> bool:
        return (word,) in self.ngrams

    def log10_probability(self, word: str, history: Sequence[str]) -> float:
        if not self.knows(word):
            word = "<unk>"
            if not self.knows(word):
                return self.unknown
        history = tuple(history[-(self.order - 1):]) if self.order > 1 else ()
        penalty = 0.0
        while True:
            entry = self.ngrams.get(history + (word,))
            if entry is not None:
                return penalty + entry[0]
            if not history:
                return penalty + self.unknown
            context = self.ngrams.get(history)
            if context is not None:
                penalty += context[1]
            history = history[1:]


class WordLanguageModelScorer:
    """The scorer hooks of TF's CTC beam search with a word n-gram model behind them.

    State = (incomplete word, history of finished words, weighted LM total of the finished words,
    score = that total + look-ahead of the incomplete word, delta = score - score of the state it was
    expanded from).  `alphabet[label]` is the character of a label; `" "` ends a word."""

    def __init__(self, model: BackOffModel, alphabet: Sequence[str], kenlm_weight: float = .8,
                 word_count_weight: float = 0., valid_word_count_weight: float = 2.3):
        self.model, self.alphabet = model, list(alphabet)
        self.weight = kenlm_weight * LN10
        self.word_count_weight, self.valid_word_count_weight = word_count_weight, valid_word_count_weight
        typeable = set(self.alphabet) - {" "}
        self.vocabulary = [w for (w,) in (k for k in model.ngrams if len(k) == 1)
                           if w not in ("<s>", "</s>", "<unk>") and w and set(w) <= typeable]
        self.min_unigram: Dict[str, float] = {}  # prefix -> lowest unigram log10 probability of a word below it
        for w in self.vocabulary:
            p = model.ngrams[(w,)][0]
            for n in range(len(w) + 1):
                prefix = w[:n]
                self.min_unigram[prefix] = min(self.min_unigram.get(prefix, 0.0), p)

    def _look_ahead(self, incomplete: str) -> float:
        if not incomplete:
            return 0.0
        return self.weight * self.min_unigram.get(incomplete, self.model.unknown)

    def _word(self, word: str, history: Tuple[str, ...]) -> Tuple[float, Tuple[str, ...]]:
        known = word in self.vocabulary_set
        total = self.weight * self.model.log10_probability(word if known else "<unk>", history)
        total += self.word_count_weight + (self.valid_word_count_weight if known else 0.0)
        return total, (history + (word if known else "<unk>",))[-max(self.model.order - 1, 1):]

    @property
    def vocabulary_set(self):
        if not hasattr(self, "_vocabulary_set"):
            self._vocabulary_set = set(self.vocabulary)
        return self._vocabulary_set

    def initialize_state(self):
        return ("", ("<s>",), 0.0, 0.0, 0.0)

    def expand_state(self, from_state, from_label: int, to_label: int):
        incomplete, history, total, score, _ = from_state
        if self.alphabet[to_label] != " ":
            incomplete = incomplete + self.alphabet[to_label]
            new_score = total + self._look_ahead(incomplete)
            return (incomplete, history, total, new_score, new_score - score)
        gained, history = self._word(incomplete, history)
        total += gained
        return ("", history, total, total, total - score)

    def expansion_score(self, state, previous: float) -> float:
        return previous + state[4]

    def expand_state_end(self, state):
        incomplete, history, total, score, _ = state
        if incomplete:
            gained, history = self._word(incomplete, history)
            total += gained
        total += self.weight * self.model.log10_probability("</s>", history)
        return ("", history, total, total, total - score)

    def end_expansion_score(self, state) -> float:
        return state[4]

    def sentence_score(self, text: str) -> float:
        """What a finished hypothesis collects in total (== NBestRescorer.score minus the CTC term)."""
        history: Tuple[str, ...] = ("<s>",)
        total = 0.0
        words = text.split(" ") if text else []
        if words and words[-1] == "":  # a trailing space has already finished the last word
            words = words[:-1]
        for word in words:
            gained, history = self._word(word, history)
            total += gained
        return total + self.weight * self.model.log10_probability("</s>", history)
