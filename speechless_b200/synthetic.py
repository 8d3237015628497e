"""Synthetic LibriSpeech-shaped batches for benchmarks, smoke and parity tests:
spectrogram values i.i.d. N(0,1) (the z-normalised contract of labeled_example.py:28-29,140),
labels uniform over the alphabet with ~15 characters per second, respecting the reference's
own feasibility rule "at least 2 prediction frames per character" (german_corpus.py:81)."""
from typing import List, Optional, Sequence

import numpy as np

from speechless_b200.labeled_example import ArrayLabeledSpectrogram

FRAMES_PER_SECOND = 16000 / 128  # sample rate / hop (labeled_example.py:77-82)


def frames_for_seconds(seconds: float) -> int:
    """STFT frame count with librosa's center=True: n // hop + 1 (10 s -> 1251)."""
    return int(seconds * 16000) // 128 + 1


def label_length_for(frames: int, chars_per_second: float = 15.0) -> int:
    seconds = (frames - 1) / FRAMES_PER_SECOND
    return max(1, min(int(chars_per_second * seconds), (frames // 2) // 2))


def random_label(rng: np.random.Generator, length: int, allowed_characters: Sequence[str]) -> str:
    return "".join(allowed_characters[i] for i in rng.integers(0, len(allowed_characters), size=length))


def synthetic_batch(batch_size: int, frames, allowed_characters: Sequence[str], feature_count: int = 128,
                    seed: int = 1234, label_length: Optional[int] = None) -> List[ArrayLabeledSpectrogram]:
    """`frames` is one length for all utterances or a per-utterance sequence."""
    rng = np.random.default_rng(seed)
    lengths = [int(frames)] * batch_size if np.isscalar(frames) else [int(f) for f in frames]
    batch = []
    for index, T in enumerate(lengths):
        spectrogram = rng.standard_normal((T, feature_count), dtype=np.float32)
        L = label_length if label_length is not None else label_length_for(T)
        batch.append(ArrayLabeledSpectrogram("synthetic-{}".format(index),
                                             random_label(rng, L, allowed_characters), spectrogram))
    return batch
