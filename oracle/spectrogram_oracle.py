"""CPU oracle for the audio front end — TEST INFRASTRUCTURE ONLY (see keras_tf_oracle.py).

Restates what `speechless/labeled_example.py` computes with third-party librosa (un-vendored,
unpinned — `requirements.txt:2`; not installable offline):

    LabeledExample.z_normalized_transposed_spectrogram()          labeled_example.py:136-140
      = z_normalize( mel( power_level( |stft|^2 ) ).T )           :99-134, 153-160, 28-29
    librosa.stft(y, n_fft=512, hop_length=128)   -> periodic Hann window, center=True with reflect
                                                    padding, frames at multiples of the hop, rfft
    power level                                   -> 10 log10(x), floored at -150 dB (0 -> -150)
    librosa.filters.mel(sr, n_fft, n_mels)        -> Slaney mel scale, triangular filters, Slaney
                                                    (area) normalisation, fmin 0, fmax sr/2
    note: the mel projection is applied to the dB values, as the reference does.

Pinning: the mel scale and filterbank are PINNED to librosa's documented values (docstring examples of
`hz_to_mel`, `mel_to_hz`, `mel_frequencies(n_mels=40)` — all 40 frequencies — and
`filters.mel(sr=22050, n_fft=2048)`; tests/test_oracle_published_vectors.py).  The whole pipeline against
librosa itself stays unpinned: the reference's only test of it (`test_labeled_example.py:14-21`) needs
librosa and downloaded audio.  Cross-validated in tests/test_spectrogram_oracle.py against independent
implementations available offline: `torch.stft` and
`transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney")`.
"""
import numpy as np


def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def frame_count(sample_count: int, hop_length: int = 128) -> int:
    """center=True: 1 + len(y) // hop frames (10 s at 16 kHz -> 1251)."""
    return 1 + sample_count // hop_length


def stft(y: np.ndarray, n_fft: int = 512, hop_length: int = 128) -> np.ndarray:
    """-> complex (1 + n_fft/2, frames)."""
    y = np.asarray(y, dtype=np.float64)
    padded = np.pad(y, n_fft // 2, mode="reflect")
    frames = frame_count(len(y), hop_length)
    window = hann_periodic(n_fft)
    out = np.empty((n_fft // 2 + 1, frames), dtype=np.complex128)
    for t in range(frames):
        out[:, t] = np.fft.rfft(window * padded[t * hop_length:t * hop_length + n_fft])
    return out


def power_to_decibel(power: np.ndarray, min_decibel: float = -150.0) -> np.ndarray:
    with np.errstate(divide="ignore"):
        level = 10.0 * np.log10(power)
    return np.where(power == 0, min_decibel, np.maximum(level, min_decibel))


def hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sample_rate: int = 16000, n_fft: int = 512, n_mels: int = 128) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels): (n_mels, 1 + n_fft/2)."""
    fft_freqs = np.linspace(0, sample_rate / 2, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sample_rate / 2), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_freqs[None, :]
    weights = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return weights * enorm[:, None]


def mel_power_level_spectrogram(y: np.ndarray, sample_rate: int = 16000, n_fft: int = 512, hop_length: int = 128,
                                n_mels: int = 128) -> np.ndarray:
    """(n_mels, frames): mel projection of the dB power spectrogram (labeled_example.py:120-134)."""
    power = np.abs(stft(y, n_fft, hop_length)) ** 2
    return mel_filterbank(sample_rate, n_fft, n_mels) @ power_to_decibel(power)


def z_normalize(a: np.ndarray) -> np.ndarray:
    return (a - np.mean(a)) / np.std(a)


def z_normalized_transposed_spectrogram(y: np.ndarray, sample_rate: int = 16000, n_fft: int = 512,
                                        hop_length: int = 128, n_mels: int = 128) -> np.ndarray:
    """(frames, n_mels), zero mean / unit variance over the whole array (labeled_example.py:136-140)."""
    return z_normalize(mel_power_level_spectrogram(y, sample_rate, n_fft, hop_length, n_mels).T)
