"""Input contract of the hot path (reference labeled_example.py:63-71): anything with an
`id`, a `label` string and `z_normalized_transposed_spectrogram() -> ndarray (T, F)`.

The reference's producers (audio -> STFT -> mel -> z-normalisation, `.npy` cache;
labeled_example.py:74-287) are outside this hot path; synthetic and
pre-computed spectrograms enter through `ArrayLabeledSpectrogram`."""
from abc import ABCMeta, abstractmethod

import numpy as np
from numpy import ndarray


def z_normalize(array: ndarray) -> ndarray:
    """Global zero-mean / unit-variance normalisation (reference labeled_example.py:28-29)."""
    return (array - np.mean(array)) / np.std(array)


class LabeledSpectrogram(metaclass=ABCMeta):
    def __init__(self, id: str, label: str):
        self.label = label
        self.id = id

    @abstractmethod
    def z_normalized_transposed_spectrogram(self) -> ndarray:
        raise NotImplementedError


class ArrayLabeledSpectrogram(LabeledSpectrogram):
    """A labeled example whose (T, F) spectrogram is already in memory."""

    def __init__(self, id: str, label: str, spectrogram: ndarray, normalize: bool = False):
        super().__init__(id=id, label=label)
        self._spectrogram = z_normalize(spectrogram) if normalize else spectrogram

    def z_normalized_transposed_spectrogram(self) -> ndarray:
        return self._spectrogram
