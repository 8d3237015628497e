"""A numpy model of the FFT of speechless_b200/csrc/frontend.cu (`spectrogram_kernel`): bit-reversed scatter
into a buffer with one pad slot per 32 elements, nine in-place radix-2 stages with per-stage contiguous
twiddle tables, and the separation of the two real frames that were packed as one complex signal.  Checked
against numpy.fft (CPU, `-m "not gpu"`); the kernel is checked on the GPU by test_gpu_spectrogram_front_end."""
import numpy as np

N = 512


def px(i):
    return i + (i >> 5)


def bit_reverse9(n):
    return int(format(n, "09b")[::-1], 2)


def paired_fft_model(frame0: np.ndarray, frame1: np.ndarray):
    x = np.zeros(N + N // 32, dtype=np.complex128)
    for n in range(N):
        x[px(bit_reverse9(n))] = frame0[n] + 1j * frame1[n]
    twiddles = np.zeros(N - 1, dtype=np.complex128)  # stage s, butterfly j at 2^s - 1 + j
    for idx in range(N - 1):
        stage = (idx + 1).bit_length() - 1
        j = idx + 1 - (1 << stage)
        twiddles[idx] = np.exp(-2j * np.pi * (j << (8 - stage)) / N)
    for stage in range(9):
        half = 1 << stage
        for tid in range(256):  # one butterfly per thread
            j = tid & (half - 1)
            base = ((tid >> stage) << (stage + 1)) + j
            w = twiddles[half - 1 + j]
            ia, ib = px(base), px(base + half)
            a, wb = x[ia], w * x[ib]
            x[ia], x[ib] = a + wb, a - wb
    spectrum0 = np.zeros(N // 2 + 1, dtype=np.complex128)
    spectrum1 = np.zeros(N // 2 + 1, dtype=np.complex128)
    for k in range(N // 2 + 1):
        z, zc = x[px(k)], x[px((N - k) & (N - 1))]
        spectrum0[k] = complex(0.5 * (z.real + zc.real), 0.5 * (z.imag - zc.imag))
        spectrum1[k] = complex(0.5 * (z.imag + zc.imag), 0.5 * (zc.real - z.real))
    return spectrum0, spectrum1


def test_paired_padded_fft_equals_numpy_rfft():
    rng = np.random.default_rng(5)
    window = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)  # periodic Hann
    for _ in range(3):
        f0, f1 = rng.standard_normal(N) * window, rng.standard_normal(N) * window
        s0, s1 = paired_fft_model(f0, f1)
        assert np.abs(s0 - np.fft.rfft(f0)).max() < 1e-9
        assert np.abs(s1 - np.fft.rfft(f1)).max() < 1e-9
    # an odd last frame is paired with zeros
    s0, s1 = paired_fft_model(f0, np.zeros(N))
    assert np.abs(s0 - np.fft.rfft(f0)).max() < 1e-9 and np.abs(s1).max() < 1e-12
    # the padded index is a bijection onto distinct slots, and a bit-reversed half-warp hits 16 distinct bank pairs
    assert len({px(i) for i in range(N)}) == N
    for first in range(0, N, 16):
        banks = {(2 * px(bit_reverse9(n))) % 32 for n in range(first, first + 16)}
        assert len(banks) == 16
