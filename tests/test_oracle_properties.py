"""Property tests of the oracle (hypothesis): size-independent facts the CUDA path is later held to
as well (tests/test_gpu_parity.py checks the same properties at BASELINE sizes)."""
from itertools import product

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import keras_tf_oracle as oracle


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 6), st.integers(2, 5), st.integers(0, 10_000))
def test_ctc_probabilities_of_all_labelings_sum_to_one(P, V, seed):
    """sum over every label sequence of p(label | x) = 1 (the collapse map is a function on paths)."""
    rng = np.random.default_rng(seed)
    lp = np.log(oracle.softmax(rng.standard_normal((P, V))))
    blank = V - 1
    total = 0.0
    for length in range(0, P + 1):
        for label in product(range(V - 1), repeat=length):
            if not oracle.ctc_feasible(list(label), P):
                continue
            _, _, ll, _ = oracle.ctc_alpha_beta(lp, list(label), blank)
            total += np.exp(ll)
    assert abs(total - 1.0) < 1e-9


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 12), st.integers(3, 8), st.integers(0, 10_000))
def test_ctc_loss_nonnegative_alpha_beta_agree_and_gradient_sums_to_zero(P, V, seed):
    rng = np.random.default_rng(seed)
    probs = oracle.softmax(rng.standard_normal((1, P, V)) * 3)
    L = int(rng.integers(0, max(1, P // 2) + 1))
    label = rng.integers(0, V - 1, size=L)
    labels = -np.ones((1, max(1, L)), dtype=np.int32)
    labels[0, :L] = label
    if not oracle.ctc_feasible(list(label), P):
        return
    losses, dlogits = oracle.ctc_batch_cost_with_logit_grad(probs, labels, [P], [L])
    assert losses[0] >= -1e-12
    lp = oracle.ctc_log_probs(probs[0])
    alpha, beta, ll, ext = oracle.ctc_alpha_beta(lp, list(label), V - 1)
    # every time slice of alpha*beta/y sums to the same likelihood
    for t in range(P):
        slice_ll = np.logaddexp.reduce(alpha[t] + beta[t] - lp[t, ext])
        assert abs(slice_ll - ll) < 1e-9
    assert np.abs(dlogits.sum(axis=2)).max() < 1e-9  # softmax Jacobian annihilates constants


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(0, 4), min_size=0, max_size=30))
def test_greedy_decode_is_idempotent_on_its_own_output(frames):
    """Decoding the one-hot rendering of a decoded sequence separated by blanks returns it unchanged."""
    V, blank = 5, 4
    probs = np.zeros((1, max(1, len(frames)), V))
    for t, c in enumerate(frames):
        probs[0, t, c] = 1
    if not frames:
        probs[0, 0, blank] = 1
    dense, lens = oracle.greedy_decode(probs, [probs.shape[1]])
    decoded = list(dense[0, :lens[0]])
    # reference rule: drop repeats, then blanks
    expected = [c for i, c in enumerate(frames) if c != blank and (i == 0 or frames[i - 1] != c)]
    assert decoded == expected
    again = np.zeros((1, max(1, 2 * len(decoded)), V))
    again[0, :, blank] = 1
    for i, c in enumerate(decoded):
        again[0, 2 * i] = 0
        again[0, 2 * i, c] = 1
    dense2, lens2 = oracle.greedy_decode(again, [again.shape[1]])
    assert list(dense2[0, :lens2[0]]) == decoded


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 40), st.integers(1, 9), st.sampled_from([1, 2]), st.integers(0, 1000))
def test_conv_same_is_linear_and_shift_consistent(T, k, stride, seed):
    rng = np.random.default_rng(seed)
    x1, x2 = rng.standard_normal((2, 1, T, 3))
    w = rng.standard_normal((k, 3, 2))
    y = oracle.conv1d_same(x1 + 2 * x2, w, None, stride)
    assert np.abs(y - (oracle.conv1d_same(x1, w, None, stride) + 2 * oracle.conv1d_same(x2, w, None, stride))).max() < 1e-10
    assert y.shape[1] == -(-T // stride)
