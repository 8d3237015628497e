"""speechless_b200 — B200-native (sm_100a) implementation of the wav2letter hot path of
juliuskunze/speechless: `Wav2Letter` (Conv1D tower + CTC + greedy decode + DP training) on
hand-written tcgen05/TMA kernels behind a C-ABI.  See DESIGN.md."""

english_frequent_characters = list("abcdefghijklmnopqrstuvwxyz '")  # reference english_corpus.py:19
german_frequent_characters = english_frequent_characters + list("äöüß")  # reference german_corpus.py:14

__all__ = ["english_frequent_characters", "german_frequent_characters"]
