"""Alphabet <-> integer grapheme mapping for the CTC (and ASG) label encodings.

Host-side mirror of the reference module of the same (mis-spelt) name
`speechless/grapheme_enconding.py` — same classes, method names, argument meaning and
exceptions, so `Wav2Letter` and the reference's tests can use it unchanged:

* `encode_label_batch`     reference :25-32   (-1 padded int32 rows)
* `decode_graphemes`       reference :34-39   (optional merge of repeats, then per-grapheme decode)
* `decode_prediction_batch`/`decode_grapheme_batch`  reference :41-57
* `CtcGraphemeEncoding`    reference :121-137 (blank is the LAST index, as tf.nn.ctc_loss needs)
* `AsgGraphemeEncoding`    reference :64-118  (two extra "repeat twice / thrice" symbols)
"""
from typing import List, Optional, Sequence

import numpy as np


def _collapse_runs(values: Sequence[int]) -> List[int]:
    """[a,a,b,b,a] -> [a,b,a]"""
    collapsed: List[int] = []
    for value in values:
        if not collapsed or collapsed[-1] != value:
            collapsed.append(value)
    return collapsed


class GraphemeEncodingBase:
    def __init__(self, allowed_characters: List[str], special_grapheme_count: int):
        self.allowed_characters = allowed_characters
        self.allowed_character_count = len(allowed_characters)
        self.grapheme_set_size = self.allowed_character_count + special_grapheme_count
        self.graphemes_by_character = {character: grapheme for grapheme, character in enumerate(allowed_characters)}

    # ---- encoding -------------------------------------------------------------------
    def encode_character(self, label_char: str) -> int:
        grapheme = self.graphemes_by_character.get(label_char)
        if grapheme is None:
            raise ValueError("Unexpected char: '{}'".format(label_char))
        return grapheme

    def encode(self, label: str) -> List[int]:
        raise NotImplementedError

    def encode_label_batch(self, labels: List[str]) -> np.ndarray:
        encoded = [self.encode(label) for label in labels]
        width = max(len(label) for label in labels)
        # the reference sizes the batch by the *string* lengths (reference :27-28)
        label_batch = np.full((len(labels), width), -1, dtype=np.int32)
        for row, (label, graphemes) in enumerate(zip(labels, encoded)):
            label_batch[row, :len(label)] = np.asarray(graphemes, dtype=np.int32)
        return label_batch

    # ---- decoding -------------------------------------------------------------------
    def decode_grapheme(self, grapheme: int, previous_grapheme: Optional[int]) -> str:
        raise NotImplementedError

    def decode_graphemes(self, graphemes: List[int], merge_repeated: bool = True) -> str:
        sequence = _collapse_runs(graphemes) if merge_repeated else list(graphemes)
        pieces = []
        previous = None
        for grapheme in sequence:
            pieces.append(self.decode_grapheme(grapheme, previous_grapheme=previous))
            previous = grapheme
        return "".join(pieces)

    def decode_grapheme_batch(self, grapheme_batch: np.ndarray, prediction_lengths: List[int],
                              merge_repeated: bool = True) -> List[str]:
        """grapheme_batch: (example, time)."""
        return [self.decode_graphemes([int(g) for g in grapheme_batch[row][:prediction_lengths[row]]],
                                      merge_repeated=merge_repeated)
                for row in range(grapheme_batch.shape[0])]

    def decode_prediction_batch(self, prediction_batch: np.ndarray, prediction_lengths: List[int]) -> List[str]:
        """prediction_batch: (example, time, grapheme) scores; greedy argmax per frame."""
        return self.decode_grapheme_batch(np.argmax(prediction_batch, axis=2), prediction_lengths)


class CtcGraphemeEncoding(GraphemeEncodingBase):
    def __init__(self, allowed_characters: List[str]):
        super().__init__(allowed_characters, special_grapheme_count=1)
        # TensorFlow's ctc_loss reserves the last class for the blank (reference :125-126)
        self.ctc_blank = self.grapheme_set_size - 1

    def encode(self, label: str) -> List[int]:
        return [self.encode_character(character) for character in label]

    def decode_grapheme(self, grapheme: int, previous_grapheme: Optional[int]) -> str:
        if 0 <= grapheme < self.allowed_character_count:
            return self.allowed_characters[grapheme]
        if grapheme == self.ctc_blank:
            return ""
        raise ValueError("Unexpected grapheme: '{}'".format(grapheme))


class AsgGraphemeEncoding(GraphemeEncodingBase):
    """ASG label encoding: runs of 2 / 3 equal characters become `char, twice` / `char, thrice`."""

    def __init__(self, allowed_characters: List[str]):
        super().__init__(allowed_characters, special_grapheme_count=2)
        self.asg_twice = self.grapheme_set_size - 2
        self.asg_thrice = self.grapheme_set_size - 1

    def encode(self, label: str) -> List[int]:
        plain = [self.encode_character(character) for character in label]
        encoded: List[int] = []
        position = 0
        while position < len(plain):
            run = 1
            while position + run < len(plain) and plain[position + run] == plain[position]:
                run += 1
            if run > 3:
                raise ValueError("{}-fold repetition found, ASG only supports up to 3-fold.".format(run))
            encoded.append(plain[position])
            if run == 2:
                encoded.append(self.asg_twice)
            elif run == 3:
                encoded.append(self.asg_thrice)
            position += run
        return encoded

    def decode_grapheme(self, grapheme: int, previous_grapheme: Optional[int]) -> str:
        if 0 <= grapheme < self.allowed_character_count:
            return self.allowed_characters[grapheme]
        if grapheme == self.asg_twice:
            return self.allowed_characters[previous_grapheme]
        if grapheme == self.asg_thrice:
            if previous_grapheme is None or not (0 <= previous_grapheme < self.allowed_character_count):
                return ""
            return self.allowed_characters[previous_grapheme] * 2
        raise ValueError("Unexpected grapheme: '{}'".format(grapheme))
