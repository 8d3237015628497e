"""A numpy model of the halo-blocked CTC lattice walk of speechless_b200/csrc/ctc.cu
(`ctc_lattice_halo_kernel`), checked against the oracle's plain alpha/beta recurrence.

It mirrors the kernel's data movement, not its speed: warps of 32 lanes x SPT states whose first 2K
slots are a halo over the previous warp's last owned states, K steps between exchanges, the finite
"log 0" sentinel with additive exists/skip biases, the reversed problem for beta, and the split of
the warps over the CTAs of a cluster.  It guards the index arithmetic on the CPU (`-m "not gpu"`);
the CUDA kernel itself is checked on the GPU by tests/test_gpu_parity.py and tools/selftest.
"""
import numpy as np
import pytest

from oracle import keras_tf_oracle as oracle

NEG = np.float64(-1e30)


def lattice_walk_model(lp, label, blank, spt, k_steps, cluster, reverse):
    """lp (P, V) natural-log probabilities -> lattice (P, S) in natural logs, like the kernel's HBM output
    (converted back from base 2), for alpha (reverse=False) or beta (reverse=True)."""
    P, _ = lp.shape
    S = 2 * len(label) + 1
    halo, slots = 2 * k_steps, 32 * spt
    own = slots - halo
    warps_all = -(-S // own)
    warps_per_cta = -(-warps_all // cluster)
    n_warps = warps_per_cta * cluster
    ext = np.full(S, blank)
    ext[1::2] = label
    if reverse:
        ext = ext[::-1]
    lp2 = (lp[::-1] if reverse else lp) * np.log2(np.e)  # walking order, base 2

    # per-slot constants of every warp: symbol, exists bias, skip bias
    state = np.arange(n_warps)[:, None] * own - halo + np.arange(slots)[None, :]
    on = (state >= 0) & (state < S)
    symbol = np.where(on, ext[np.clip(state, 0, S - 1)], blank)
    prev2 = ext[np.clip(state - 2, 0, S - 1)]
    skip = on & (state % 2 == 1) & (state >= 3) & (symbol != prev2)
    on_bias = np.where(on, 0.0, NEG)
    skip_bias = np.where(skip, 0.0, NEG)
    assert ((state % 2 == 0) | ~on | (symbol != blank)).all()  # odd slots carry labels, even slots blanks

    # published columns: one buffer per CTA (index = local state + halo), double buffered by block parity
    stride = warps_per_cta * own + halo
    column = np.full((cluster, 2, stride), NEG)
    column[0, 1, halo] = 0.0  # virtual column of step -1: state 0 holds log 1
    out = np.full((P, S), -np.inf)
    for t0 in range(0, P, k_steps):
        kb = t0 // k_steps
        regs = np.empty((n_warps, slots))
        for w in range(n_warps):  # reload halo + own states from the CTA's buffer of the previous block
            cta, lw = divmod(w, warps_per_cta)
            regs[w] = column[cta, (kb + 1) & 1, lw * own:lw * own + slots]
        for t in range(t0, min(t0 + k_steps, P)):
            left = np.empty_like(regs)  # value of the slot to the left: shuffle, lane 0 keeps its own value
            left[:, 1:] = regs[:, :-1]
            left[:, 0] = regs[:, spt - 1]
            left2 = np.empty_like(regs)
            left2[:, 2:] = regs[:, :-2]
            left2[:, :2] = regs[:, :2]  # (inside the halo: garbage by construction, never reaches an owned slot)
            a2 = left2 + skip_bias
            m = np.maximum(np.maximum(regs, left), a2)
            total = np.exp2(regs - m) + np.exp2(left - m) + np.exp2(a2 - m)
            regs = m + np.log2(total) + lp2[t][symbol] + on_bias
            for w in range(n_warps):  # owned, existing states go to HBM
                sel = on[w, halo:]
                out[t, state[w, halo:][sel]] = regs[w, halo:][sel]
        for w in range(n_warps):  # publish the owned states; the last warp of a CTA also pushes its last
            cta, lw = divmod(w, warps_per_cta)  # 2K owned states into the next CTA's halo slots (DSMEM)
            column[cta, kb & 1, lw * own + halo:lw * own + slots] = regs[w, halo:]
            if lw == warps_per_cta - 1 and cta + 1 < cluster:
                column[cta + 1, kb & 1, :halo] = regs[w, slots - halo:]
    out = out / np.log2(np.e)
    out[out < -1e29] = -np.inf
    return out[::-1, ::-1] if reverse else out


@pytest.mark.parametrize("spt,k_steps,cluster", [(2, 8, 1), (2, 4, 1), (4, 8, 1), (4, 32, 4), (2, 16, 2), (8, 16, 8)])
def test_halo_blocked_walk_equals_the_plain_recurrence(spt, k_steps, cluster):
    rng = np.random.default_rng(spt * 100 + k_steps + cluster)
    V, blank = 7, 6
    for P, L in ((41, 17), (70, 30), (9, 0), (33, 16)):
        label = rng.integers(0, V - 1, size=L)
        if L > 4:
            label[2] = label[1]  # a repeated character (no skip transition across it)
        probs = oracle.softmax(rng.standard_normal((P, V)) * 3)
        lp = oracle.ctc_log_probs(probs)
        alpha, beta, ll, _ = oracle.ctc_alpha_beta(lp, label, blank)
        got_alpha = lattice_walk_model(lp, label, blank, spt, k_steps, cluster, reverse=False)
        got_beta = lattice_walk_model(lp, label, blank, spt, k_steps, cluster, reverse=True)
        for got, want in ((got_alpha, alpha), (got_beta, beta)):
            finite = np.isfinite(want)
            assert (np.isfinite(got) == finite).all()
            assert np.abs(got[finite] - want[finite]).max() < 1e-9
        S = 2 * L + 1
        last = got_alpha[P - 1]
        total = last[S - 1] if S == 1 else np.logaddexp(last[S - 1], last[S - 2])
        assert total == pytest.approx(ll, abs=1e-9)
