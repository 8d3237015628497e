"""A numpy model of the candidate selection of speechless_b200/csrc/beam.cu (`ctc_beam_search_kernel`,
step 4): order-preserving integer image of the fp32 scores, bit-wise binary search for the score of the
W-th best candidate with early exit, ties at the threshold resolved towards the lower candidate index.
Checked against a plain sort (CPU, `-m "not gpu"`); the kernel itself is checked on the GPU by
tests/test_beam_search.py."""
import numpy as np
import pytest

ORD_NEG_INF = 0x007FFFFF


def float_to_ordered(values: np.ndarray) -> np.ndarray:
    bits = values.astype(np.float32).view(np.uint32).astype(np.uint64)
    negative = (bits & 0x80000000) != 0
    return np.where(negative, (~bits) & 0xFFFFFFFF, bits | 0x80000000).astype(np.uint64)


def select_model(scores: np.ndarray, beam_width: int):
    """-> (selected candidate indices in slot order, number of block-wide counts used)."""
    keys = float_to_ordered(scores)
    counts = 1
    n_next = min(beam_width, int((keys > ORD_NEG_INF).sum()))
    thr, exact = 0, False
    for bit in range(31, -1, -1):
        if exact:
            break
        candidate = thr | (1 << bit)
        counts += 1
        c = int((keys >= candidate).sum())
        if c >= n_next:
            thr = candidate
        exact = c == n_next
    index = np.arange(keys.size)
    idx_limit = keys.size
    if exact:
        thr -= 1
        idx_limit = -1
    else:
        counts += 2
        n_greater, n_ties = int((keys > thr).sum()), int((keys == thr).sum())
        if n_greater + n_ties > n_next:
            need, lo = n_next - n_greater, 0
            for bit in range(11, -1, -1):
                candidate = lo | (1 << bit)
                counts += 1
                if int(((keys == thr) & (index < candidate)).sum()) < need:
                    lo = candidate
            idx_limit = lo
    chosen = (keys > thr) | ((keys == thr) & (keys > ORD_NEG_INF) & (index <= idx_limit))
    packed = (keys[chosen] << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - index[chosen].astype(np.uint64))
    order = np.argsort(packed)[::-1]  # rank by counting == descending order of the unique 64-bit keys
    return index[chosen][order], counts


def reference_selection(scores: np.ndarray, beam_width: int):
    finite = np.flatnonzero(np.isfinite(scores.astype(np.float32)) | (scores.astype(np.float32) > -np.inf))
    finite = [i for i in finite if scores[i] > -np.inf]
    ranked = sorted(finite, key=lambda i: (-np.float32(scores[i]), i))
    return ranked[:beam_width]


@pytest.mark.parametrize("case", ["random", "ties", "few_finite", "all_equal", "mixed_signs"])
def test_threshold_search_selects_the_top_candidates(case):
    rng = np.random.default_rng(hash(case) % 1000)
    for _ in range(40):
        n, width = int(rng.integers(1, 400)), int(rng.integers(1, 130))
        scores = rng.normal(size=n) * 20 - 30
        if case == "ties":
            scores = np.round(scores / 4) * 4  # many exactly equal scores, also at the threshold
        elif case == "few_finite":
            scores[rng.random(n) < 0.9] = -np.inf
            scores[int(rng.integers(0, n))] = -3.0
        elif case == "all_equal":
            scores[:] = -7.25
        elif case == "mixed_signs":
            scores = rng.normal(size=n) * 5  # positive log-scores never occur, the mapping must still order them
        scores = scores.astype(np.float32) + np.float32(0.0)  # (-0.0 -> +0.0: the integer image orders the two zeros)
        got, counts = select_model(scores, width)
        assert list(got) == reference_selection(scores, width)
        assert counts <= 1 + 32 + 2 + 12
