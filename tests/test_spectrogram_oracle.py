"""Cross-validation of oracle/spectrogram_oracle.py (librosa is not installable offline)."""
import numpy as np
import pytest
import torch

from oracle import spectrogram_oracle as so


def _signal(seconds=0.7, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(int(16000 * seconds)) / 16000
    return (0.4 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 3000 * t * (1 + 0.2 * t)) +
            0.05 * rng.standard_normal(t.shape))


def test_stft_matches_torch_stft():
    y = _signal()
    ours = so.stft(y)
    ref = torch.stft(torch.tensor(y), n_fft=512, hop_length=128, window=torch.hann_window(512, periodic=True,
                     dtype=torch.float64), center=True, pad_mode="reflect", return_complex=True).numpy()
    assert ours.shape == ref.shape == (257, so.frame_count(len(y)))
    assert np.abs(ours - ref).max() < 1e-9
    assert so.frame_count(160000) == 1251  # 10 s -> the tower's T


def test_mel_filterbank_matches_transformers_slaney():
    audio_utils = pytest.importorskip("transformers.audio_utils")
    ref = audio_utils.mel_filter_bank(num_frequency_bins=257, num_mel_filters=128, min_frequency=0.0,
                                      max_frequency=8000.0, sampling_rate=16000, norm="slaney", mel_scale="slaney")
    ours = so.mel_filterbank(16000, 512, 128)
    assert ours.shape == (128, 257)
    assert np.abs(ours - np.asarray(ref).T).max() < 1e-9
    # Slaney scale anchors: linear below 1 kHz (200/3 Hz per mel), log above
    assert abs(float(so.hz_to_mel(1000.0)) - 15.0) < 1e-12
    assert abs(float(so.mel_to_hz(so.hz_to_mel(4321.0))) - 4321.0) < 1e-9


def test_power_level_floor_and_normalisation():
    assert so.power_to_decibel(np.array([0.0, 1.0, 1e-20, 100.0])).tolist() == [-150.0, 0.0, -150.0, 20.0]
    z = so.z_normalized_transposed_spectrogram(_signal(0.3, seed=2))
    assert z.shape == (so.frame_count(4800), 128)
    assert abs(z.mean()) < 1e-12 and abs(z.std() - 1) < 1e-12
