"""Cross-validates the numpy oracle against independent implementations available offline
(torch.nn.functional.conv1d / ctc_loss / autograd, finite differences, brute-force path
enumeration).  Keras/TF themselves cannot run here, so for the Conv1D tower and Adam this is
the strongest available check ("parity unpinned", see oracle/keras_tf_oracle.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import keras_tf_oracle as oracle
from oracle.torch_cpu import TorchCpuWav2Letter


def _torch_conv_same(x, w, b, stride):
    T = x.shape[1]
    _, pad_l, pad_r = oracle.same_padding(T, w.shape[0], stride)
    xt = F.pad(torch.as_tensor(x).transpose(1, 2), (pad_l, pad_r))
    y = F.conv1d(xt, torch.as_tensor(w).permute(2, 1, 0), torch.as_tensor(b), stride=stride)
    return y.transpose(1, 2).numpy()


@pytest.mark.parametrize("T,k,stride", [(37, 48, 2), (38, 48, 2), (50, 7, 1), (41, 32, 1), (9, 1, 1), (5, 7, 1),
                                         (1000, 250, 160), (961, 250, 160), (90, 250, 160)])  # + wave_conv (net.py:310-312)
def test_conv1d_same_matches_torch(T, k, stride):
    rng = np.random.default_rng(T * 100 + k)
    x = rng.standard_normal((2, T, 5))
    w = rng.standard_normal((k, 5, 4))
    b = rng.standard_normal(4)
    assert np.abs(oracle.conv1d_same(x, w, b, stride) - _torch_conv_same(x, w, b, stride)).max() < 1e-12


def test_conv1d_backward_matches_autograd():
    rng = np.random.default_rng(5)
    for (T, k, stride) in [(21, 48, 2), (20, 7, 1), (13, 32, 1), (700, 250, 160)]:
        x = rng.standard_normal((2, T, 3))
        w = rng.standard_normal((k, 3, 4))
        b = rng.standard_normal(4)
        dy = rng.standard_normal(oracle.conv1d_same(x, w, b, stride).shape)
        dx, dw, db = oracle.conv1d_same_backward(x, w, dy, stride)
        xt = torch.tensor(x, requires_grad=True)
        wt = torch.tensor(w, requires_grad=True)
        bt = torch.tensor(b, requires_grad=True)
        _, pad_l, pad_r = oracle.same_padding(T, k, stride)
        y = F.conv1d(F.pad(xt.transpose(1, 2), (pad_l, pad_r)), wt.permute(2, 1, 0), bt, stride=stride).transpose(1, 2)
        (y * torch.tensor(dy)).sum().backward()
        assert np.abs(dx - xt.grad.numpy()).max() < 1e-11
        assert np.abs(dw - wt.grad.numpy()).max() < 1e-11
        assert np.abs(db - bt.grad.numpy()).max() < 1e-11


def _random_ctc_case(rng, B, T, V, max_len):
    probs = oracle.softmax(rng.standard_normal((B, T, V)) * 2)
    label_lengths = rng.integers(0, max_len + 1, size=B)
    labels = -np.ones((B, max(1, label_lengths.max())), dtype=np.int32)
    pred = np.zeros(B, dtype=np.int64)
    for b in range(B):
        lab = rng.integers(0, V - 1, size=label_lengths[b])
        if label_lengths[b] > 2:
            lab[1] = lab[0]  # force a repeat
        labels[b, :label_lengths[b]] = lab
        pred[b] = rng.integers(max(1, len(lab) + sum(lab[:-1] == lab[1:])), T + 1)
    return probs, labels, pred, label_lengths


def test_ctc_matches_torch_ctc_loss_and_autograd():
    rng = np.random.default_rng(11)
    probs, labels, pred, ll = _random_ctc_case(rng, B=5, T=30, V=7, max_len=8)
    losses, dlogits = oracle.ctc_batch_cost_with_logit_grad(probs, labels, pred, ll)
    # independent: torch softmax -> log(p+eps) -> log_softmax -> ctc_loss, gradient by autograd wrt logits
    logits = torch.tensor(np.log(probs), requires_grad=True)  # softmax(log p) == p
    p = torch.softmax(logits, dim=2)
    lp = torch.log_softmax(torch.log(p + oracle.EPSILON), dim=2).transpose(0, 1)
    tl = F.ctc_loss(lp, torch.tensor(np.maximum(labels, 0), dtype=torch.long), torch.tensor(pred),
                    torch.tensor(ll), blank=6, reduction="none")
    tl.sum().backward()
    assert np.abs(losses - tl.detach().numpy()).max() < 1e-9
    assert np.abs(dlogits - logits.grad.numpy()).max() < 1e-9


def test_ctc_brute_force_small():
    rng = np.random.default_rng(3)
    for label in ([0, 0, 2], [1], [], [2, 1]):
        lp = np.log(oracle.softmax(rng.standard_normal((5, 4))))
        _, _, ll, _ = oracle.ctc_alpha_beta(lp, label, blank=3)
        assert abs(ll - oracle.ctc_brute_force_log_likelihood(lp, label, blank=3)) < 1e-10
    # value recorded during the survey: T=5, V=4, label [0,0,2]


def test_ctc_infeasible_raises():
    probs = oracle.softmax(np.zeros((1, 3, 4)))
    with pytest.raises(ValueError, match="Not enough time"):
        oracle.ctc_batch_cost(probs, np.array([[0, 0, 1]]), [3], [3])  # needs 4 frames


def test_tower_and_gradients_match_torch_cpu_variant():
    rng = np.random.default_rng(8)
    V = 6
    ref = oracle.Wav2LetterOracle(8, V, main_filter_count=6, out_filter_count=10, seed=1, dtype=np.float64)
    for b in ref.biases:  # non-zero biases: exercises the unmasked-padding bleed
        b += rng.standard_normal(b.shape) * 0.1
    fast = TorchCpuWav2Letter(8, V, 6, 10, dtype=torch.float64)
    fast.set_weights(ref.weights, ref.biases)
    x = rng.standard_normal((3, 61, 8))
    x[1, 40:] = 0  # a shorter utterance, zero padded
    labels = np.array([[0, 1, 1, 4], [2, 3, -1, -1], [4, -1, -1, -1]], dtype=np.int32)
    pred, ll = [30, 20, 30], [4, 2, 1]
    losses, probs, logits, dws, dbs = ref.loss_and_gradients(x, labels, pred, ll)
    t_probs, t_logits = fast.forward(torch.as_tensor(x), return_logits=True)
    assert np.abs(probs - t_probs.detach().numpy()).max() < 1e-12
    assert np.abs(logits - t_logits.detach().numpy()).max() < 1e-11
    t_losses, t_dws, t_dbs = fast.gradients(x, labels, pred, ll)
    assert np.abs(losses - t_losses).max() < 1e-9
    for a, b in zip(dws + dbs, t_dws + t_dbs):
        assert np.abs(a - b).max() < 1e-9 * max(1.0, np.abs(b).max())


def test_keras_adam_matches_manual_and_torch_cpu_variant():
    rng = np.random.default_rng(2)
    p0 = [rng.standard_normal((3, 4)), rng.standard_normal(5)]
    adam = oracle.KerasAdam(lr=1e-2)
    p = [a.copy() for a in p0]
    m = [np.zeros_like(a) for a in p0]
    v = [np.zeros_like(a) for a in p0]
    q = [a.copy() for a in p0]
    for t in range(1, 6):
        grads = [rng.standard_normal(a.shape) for a in p0]
        p = adam.step(p, grads)
        lr_t = 1e-2 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        for i, g in enumerate(grads):
            m[i] = 0.9 * m[i] + 0.1 * g
            v[i] = 0.999 * v[i] + 0.001 * g * g
            q[i] = q[i] - lr_t * m[i] / (np.sqrt(v[i]) + 1e-8)
    for a, b in zip(p, q):
        assert np.abs(a - b).max() < 1e-15
    # epsilon placement differs from torch.optim.Adam: make sure we are NOT accidentally torch's rule
    w = torch.tensor(p0[1].copy(), requires_grad=True)
    opt = torch.optim.Adam([w], lr=1e-2, eps=1e-3)
    w.grad = torch.tensor(np.full(5, 1e-3))
    opt.step()
    keras = oracle.KerasAdam(lr=1e-2, epsilon=1e-3).step([p0[1].copy()], [np.full(5, 1e-3)])[0]
    assert np.abs(keras - w.detach().numpy()).max() > 1e-4


def test_padding_bleed_is_reproduced():
    """padded frames are not masked; with non-zero biases an utterance's last
    valid output frames depend on how long the batch is padded."""
    rng = np.random.default_rng(4)
    ref = oracle.Wav2LetterOracle(4, 5, main_filter_count=3, out_filter_count=4, seed=2)
    for b in ref.biases:
        b += 0.3
    utt = rng.standard_normal((100, 4))
    alone = ref.forward(utt[None])[0]
    padded = np.zeros((1, 240, 4))
    padded[0, :100] = utt
    embedded = ref.forward(padded)[0, :50]
    changed = np.where(np.abs(alone - embedded).max(axis=1) > 1e-12)[0]
    assert len(changed) > 0 and changed.min() >= 50 - 37  # only the last <= 37 valid frames move


def test_raw_wave_tower_numpy_matches_torch_restatement():
    """use_raw_wave_input (net.py:310-316): 12 layers, ratio 320; numpy and torch restatements agree."""
    specs = oracle.wav2letter_layer_specs(1, 29, 16, 24, use_raw_wave_input=True)
    assert [s[0] for s in specs][:2] == ["wave_conv", "striding_conv"]
    assert specs[0][1:5] == (1, 16, 250, 160) and specs[1][1] == 16
    assert int(np.prod([s[4] for s in specs])) == 320
    ref = oracle.Wav2LetterOracle(1, 29, 16, 24, dtype=np.float64, use_raw_wave_input=True)
    cpu = TorchCpuWav2Letter(1, 29, 16, 24, dtype=torch.float64, use_raw_wave_input=True)
    cpu.set_weights(ref.weights, ref.biases)
    x = np.random.default_rng(1).standard_normal((2, 2500, 1))
    probs = ref.forward(x)
    assert probs.shape == (2, 8, 29)  # ceil(ceil(2500/160)/2)
    assert np.abs(probs - cpu.forward(torch.as_tensor(x)).detach().numpy()).max() < 1e-12
