"""Input contract of the hot path (reference labeled_example.py:63-71): anything with an
`id`, a `label` string and `z_normalized_transposed_spectrogram() -> ndarray (T, F)`.

Also the producers next to the path: `LabeledExample` (audio -> spectrogram,
reference labeled_example.py:74-171, computed by the GPU front end of `frontend.py` instead of
librosa) and `CachedLabeledSpectrogram` (the `.npy` cache, reference :236-261).  Synthetic and
pre-computed spectrograms enter through `ArrayLabeledSpectrogram`."""
from abc import ABCMeta, abstractmethod
from pathlib import Path
from typing import Callable, Optional

import numpy as np
from numpy import ndarray

from speechless_b200.tools import log, mkdir


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


class LabeledExample(LabeledSpectrogram):
    """A labeled audio clip whose spectrogram is computed on the GPU (same constructor arguments and
    defaults as the reference's class, labeled_example.py:74-92; only the reference defaults
    512 / 128 / 128 are built)."""
    _front_ends = {}

    def __init__(self, get_raw_audio: Callable[[], ndarray], sample_rate: int = 16000, id: Optional[str] = None,
                 label: Optional[str] = "nolabel", fourier_window_length: int = 512, hop_length: int = 128,
                 mel_frequency_count: int = 128):
        super().__init__(id=id, label=label)
        self.get_raw_audio = get_raw_audio
        self.sample_rate = sample_rate
        self.fourier_window_length = fourier_window_length
        self.hop_length = hop_length
        self.mel_frequency_count = mel_frequency_count

    def _front_end(self):
        from speechless_b200.frontend import SpectrogramFrontEnd
        key = (self.sample_rate, self.fourier_window_length, self.hop_length, self.mel_frequency_count)
        if key not in LabeledExample._front_ends:
            LabeledExample._front_ends[key] = SpectrogramFrontEnd(*key)
        return LabeledExample._front_ends[key]

    def z_normalized_transposed_spectrogram(self) -> ndarray:
        """(time, frequencies), zero mean and unit variance over the whole array."""
        return self._front_end().z_normalized_transposed_spectrograms([self.get_raw_audio()])[0]

    @property
    def duration_in_s(self) -> float:
        return len(self.get_raw_audio()) / self.sample_rate

    def __str__(self) -> str:
        return self.id + (": {}".format(self.label) if self.label else "")


class CachedLabeledSpectrogram(LabeledSpectrogram):
    """`.npy` cache in front of any LabeledSpectrogram: `{directory}/{id}.npy`, recomputed when the
    file is missing or unreadable (reference labeled_example.py:236-261)."""

    def __init__(self, original: LabeledSpectrogram, spectrogram_cache_directory: Path):
        super().__init__(id=original.id, label=original.label)
        self.original = original
        self.spectrogram_cache_file = Path(spectrogram_cache_directory) / "{}.npy".format(original.id)

    def is_cached(self) -> bool:
        return self.spectrogram_cache_file.exists()

    def z_normalized_transposed_spectrogram(self) -> ndarray:
        if self.is_cached():
            try:
                return np.load(str(self.spectrogram_cache_file))
            except ValueError:
                log("Recalculating cached file {} because loading failed.".format(self.spectrogram_cache_file))
        spectrogram = self.original.z_normalized_transposed_spectrogram()
        mkdir(self.spectrogram_cache_file.parent)
        np.save(str(self.spectrogram_cache_file), spectrogram)
        return spectrogram
