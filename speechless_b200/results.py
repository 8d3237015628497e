"""Result types returned by the `Wav2Letter` surface (reference net.py:22-114).

Same class names, constructor arguments, attributes and `__str__` formats as the reference
so callers such as `Configuration.test_model` / `main.py` read them unchanged.  The
reference computes error counts with the third-party `editdistance` C++ extension
(net.py:4,33,37; not installed here); `levenshtein` restates it."""
from typing import Dict, List, Sequence

from speechless_b200.tools import average_or_nan


def levenshtein(a: Sequence, b: Sequence) -> int:
    """Edit distance (insert / delete / substitute, unit costs) between two sequences —
    what `editdistance.eval` returns for strings and for lists of words."""
    if len(a) < len(b):
        a, b = b, a
    if len(b) == 0:
        return len(a)
    previous = list(range(len(b) + 1))
    for i, item_a in enumerate(a, start=1):
        current = [i] + [0] * len(b)
        for j, item_b in enumerate(b, start=1):
            substitution = previous[j - 1] + (0 if item_a == item_b else 1)
            current[j] = min(previous[j] + 1, current[j - 1] + 1, substitution)
        previous = current
    return previous[-1]


class ExpectationVsPrediction:
    def __init__(self, expected: str, predicted: str, loss: float):
        self.loss = loss
        self.expected = expected
        self.predicted = predicted
        self.expected_letter_count = len(expected)
        self.expected_words = expected.split()
        self.expected_word_count = len(self.expected_words)
        self._letter_error_count = None
        self._word_error_count = None

    @property
    def letter_error_count(self) -> float:
        if self._letter_error_count is None:
            self._letter_error_count = levenshtein(self.expected, self.predicted)
        return self._letter_error_count

    @property
    def word_error_count(self) -> float:
        if self._word_error_count is None:
            self._word_error_count = levenshtein(self.expected_words, self.predicted.split())
        return self._word_error_count

    @property
    def letter_error_rate(self) -> float:
        return self.letter_error_count / self.expected_letter_count

    @property
    def word_error_rate(self) -> float:
        return self.word_error_count / self.expected_word_count

    def __str__(self):
        return 'Expected:  "{}"\nPredicted: "{}"\nErrors: {} letters ({}%), {} words ({}%), loss: {:.2f}.'.format(
            self.expected, self.predicted,
            self.letter_error_count, round(self.letter_error_rate * 100),
            self.word_error_count, round(self.word_error_rate * 100),
            self.loss)


class ExpectationsVsPredictions:
    def __init__(self, results: List[ExpectationVsPrediction]):
        self.results = results

    def _average(self, attribute: str) -> float:
        return average_or_nan([getattr(result, attribute) for result in self.results])

    @property
    def average_letter_error_count(self) -> float:
        return self._average("letter_error_count")

    @property
    def average_word_error_count(self) -> float:
        return self._average("word_error_count")

    @property
    def average_letter_error_rate(self) -> float:
        return self._average("letter_error_rate")

    @property
    def average_word_error_rate(self) -> float:
        return self._average("word_error_rate")

    @property
    def average_loss(self) -> float:
        return self._average("loss")

    def summary_line(self) -> str:
        return ("Average over {} examples: {:.1f} letter errors ({:.2f}%), "
                "{:.1f} word errors ({:.2f}%), loss {:.2f}.").format(
            len(self.results),
            self.average_letter_error_count, self.average_letter_error_rate * 100,
            self.average_word_error_count, self.average_word_error_rate * 100,
            self.average_loss)

    def __str__(self):
        return "\n\n".join(str(result) for result in self.results) + "\n\n" + self.summary_line() + "\n\n"


class ExpectationsVsPredictionsInBatches(ExpectationsVsPredictions):
    def __init__(self, result_batches: List[ExpectationsVsPredictions]):
        self.result_batches = result_batches
        super().__init__([result for batch in result_batches for result in batch.results])

    def __str__(self):
        return "All batches: {}".format(self.summary_line())


class ExpectationsVsPredictionsInGroupedBatches(ExpectationsVsPredictions):
    def __init__(self, results_by_group_name: Dict[str, ExpectationsVsPredictionsInBatches]):
        self.result_batches_by_group_name = results_by_group_name
        super().__init__([result for batches in results_by_group_name.values() for result in batches.results])

    def __str__(self):
        groups_summary = "\n".join("{}: {}".format(group_name, batches)
                                   for group_name, batches in self.result_batches_by_group_name.items())
        return "\n\n{}\n\nAll corpora: {}\n\n".format(groups_summary, self.summary_line())
