"""Word n-gram language model for re-scoring beam-search hypotheses (host side).

The reference decodes with a language model only through a patched TensorFlow
(`tf.nn.ctc_beam_search_decoder(kenlm_directory_path=..., kenlm_weight=.8, word_count_weight=0,
valid_word_count_weight=2.3)`, reference net.py:444-451; fork named in net.py:420-422 and
README.md:17).  Neither that fork nor the KenLM library is in the reference tree or installable
here, so its in-search scorer cannot be restated; what is kept is the interface
(`kenlm_directory` with its `vocabulary` file, net.py:171-177) and the three weights, applied as
N-BEST RE-SCORING of the device beam search (`sl_ctc_beam_search_decode`):

    score(hypothesis) = log P_ctc + kenlm_weight * ln P_lm(words)
                        + word_count_weight * #words + valid_word_count_weight * #words known to the LM

**Parity unpinned** against the fork (stated in DESIGN.md).  The model file is the plain-text ARPA
format every KenLM installation can write (`lmplz`); KenLM's binary format is not read.
"""
import math
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

LN10 = math.log(10.0)


class ArpaLanguageModel:
    """Back-off n-gram model read from an ARPA file (log10 probabilities and back-off weights)."""

    def __init__(self, ngrams: Dict[Tuple[str, ...], Tuple[float, float]], order: int):
        self.ngrams = ngrams
        self.order = order
        self.unknown = ngrams.get(("<unk>",), (-100.0, 0.0))[0]

    @staticmethod
    def read(path: Path) -> "ArpaLanguageModel":
        ngrams: Dict[Tuple[str, ...], Tuple[float, float]] = {}
        order, current = 0, 0
        with open(str(path), encoding="utf8") as lines:
            for raw in lines:
                line = raw.strip()
                if not line or line == "\\data\\" or line.startswith("ngram "):
                    continue
                if line == "\\end\\":
                    break
                if line.startswith("\\") and line.endswith("-grams:"):
                    current = int(line[1:line.index("-")])
                    order = max(order, current)
                    continue
                fields = line.split("\t") if "\t" in line else line.split()
                if current == 0 or len(fields) < 2:
                    raise ValueError("Malformed ARPA line: {!r}".format(raw))
                if "\t" in line:
                    words = tuple(fields[1].split(" "))
                    backoff = float(fields[2]) if len(fields) > 2 else 0.0
                else:
                    words = tuple(fields[1:1 + current])
                    backoff = float(fields[1 + current]) if len(fields) > 1 + current else 0.0
                if len(words) != current:
                    raise ValueError("Malformed ARPA line: {!r}".format(raw))
                ngrams[words] = (float(fields[0]), backoff)
        if order == 0:
            raise ValueError("No n-grams found in {}".format(path))
        return ArpaLanguageModel(ngrams, order)

    def knows(self, word: str) -> bool:
        return (word,) in self.ngrams

    def log10_probability(self, word: str, history: Sequence[str]) -> float:
        """log10 P(word | history) with Katz back-off (the ARPA semantics KenLM implements)."""
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

    def log10_sentence(self, words: Sequence[str]) -> float:
        """log10 P(<s> words </s>) (without the probability of <s> itself, as KenLM's `score`)."""
        history: List[str] = ["<s>"]
        total = 0.0
        for word in list(words) + ["</s>"]:
            total += self.log10_probability(word, history)
            history.append(word if self.knows(word) else "<unk>")
        return total


class NBestRescorer:
    """Picks the best of the beam-search hypotheses under CTC + language-model score (weights as in
    reference net.py:448-451)."""

    def __init__(self, language_model: ArpaLanguageModel, kenlm_weight: float = .8, word_count_weight: float = 0.,
                 valid_word_count_weight: float = 2.3):
        self.language_model = language_model
        self.kenlm_weight = kenlm_weight
        self.word_count_weight = word_count_weight
        self.valid_word_count_weight = valid_word_count_weight

    def score(self, text: str, ctc_log_probability: float) -> float:
        words = text.split()
        valid = sum(1 for w in words if self.language_model.knows(w))
        return (ctc_log_probability + self.kenlm_weight * LN10 * self.language_model.log10_sentence(words) +
                self.word_count_weight * len(words) + self.valid_word_count_weight * valid)

    def best(self, hypotheses: Sequence[Tuple[str, float]]) -> Tuple[str, float]:
        """hypotheses: (text, CTC log-probability), any order; ties keep the earlier (better CTC) one."""
        scored = [(self.score(text, log_probability), -index, text) for index, (text, log_probability) in
                  enumerate(hypotheses)]
        best_score, _, best_text = max(scored)
        return best_text, best_score


def find_arpa_file(kenlm_directory: Path) -> Optional[Path]:
    candidates = sorted(Path(kenlm_directory).glob("*.arpa"))
    return candidates[0] if candidates else None
