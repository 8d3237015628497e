"""CPU oracle for the wav2letter hot path of juliuskunze/speechless — TEST INFRASTRUCTURE ONLY.

Nothing in the product path (`speechless_b200/`) imports this module; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu-baseline / `--impl reference` legs may.

What it restates
----------------
The reference executes this path inside third-party Keras 2.0.x / TensorFlow 1.x
(unpinned: `requirements.txt:3`, `README.md:11`), neither of which is vendored or
installable offline.  This file restates, in plain numpy, the published semantics of
exactly the calls the reference makes:

* `conv1d_same`          <- keras.layers.Conv1D(padding="same")      net.py:304-305
* `Wav2LetterOracle`     <- Wav2Letter.create_predictive_net         net.py:291-341
* `ctc_batch_cost`       <- K.ctc_batch_cost -> tf.nn.ctc_loss       net.py:402-406
* `greedy_decode`        <- tf.nn.ctc_greedy_decoder + post-process  net.py:453-454,468-475
* `KerasAdam`            <- keras.optimizers.Adam(1e-4)              net.py:132,389
* `input_batch_and_prediction_lengths`, `encode_label_batch`         net.py:578-587, grapheme_enconding.py:25-32

Pinning status
--------------------------------------
* greedy decode, grapheme encode/decode: PINNED by the reference's own test vectors
  (`speechless/test/test_ctc_decoders.py:22-24,38-41`, `test_grapheme_encoding.py:12-31`),
  checked in tests/test_oracle_golden.py.
* CTC loss: pinned to TensorFlow's own known-answer vectors (ctc_loss_op_test.py
  `testBasic`: -log p = 3.34211 / 5.42262) plus brute-force path enumeration; gradients
  against finite differences and torch autograd.
* Conv1D forward / input gradient / filter gradient (SAME padding incl. the asymmetric stride-2 case)
  and the Adam update: PINNED to TensorFlow's published known answers — the literal expected vectors of
  `conv_ops_test.py::Conv2DTest` (testConv2D1x1Filter, 1x2Filter, 2x2Filter, 2x2FilterStride2[Same],
  2x2Depth{1,3}ValidBackprop{Input,Filter}) and `adam_test.py::adam_update_numpy`, typed into
  tests/test_oracle_published_vectors.py (Keras' Conv1D is TF's conv2d on a height-1 image).  The
  composition of the pieces into the 11-layer tower has no published vector; it is cross-validated
  against independent implementations (torch.nn.functional.conv1d with explicit asymmetric padding,
  torch autograd, a hand-stepped Adam) in tests/test_oracle_crosscheck.py.
"""
from itertools import groupby, product
from typing import List, Optional, Sequence, Tuple

import numpy as np

EPSILON = 1e-8  # keras.backend.epsilon(), added inside K.ctc_batch_cost (tf.log(y_pred + eps))


# --------------------------------------------------------------------------------------
# Conv1D, padding="same" (TF SAME rule), channels-last, cross-correlation, kernel (k,Cin,Cout)
# --------------------------------------------------------------------------------------
def same_padding(T: int, k: int, stride: int) -> Tuple[int, int, int]:
    """TF SAME: out = ceil(T/s); total = max((out-1)s + k - T, 0); left = total//2."""
    t_out = -(-T // stride)
    total = max((t_out - 1) * stride + k - T, 0)
    pad_l = total // 2
    return t_out, pad_l, total - pad_l


def conv1d_same(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray], stride: int = 1) -> np.ndarray:
    """y[b,t,co] = bias[co] + sum_j sum_ci xpad[b, t*s + j, ci] * w[j, ci, co]  (net.py:304-305)."""
    B, T, Cin = x.shape
    k, cin2, Cout = w.shape
    assert cin2 == Cin
    t_out, pad_l, pad_r = same_padding(T, k, stride)
    xp = np.zeros((B, T + pad_l + pad_r, Cin), dtype=x.dtype)
    xp[:, pad_l:pad_l + T] = x
    y = np.zeros((B, t_out, Cout), dtype=x.dtype)
    for j in range(k):
        rows = xp[:, j:j + (t_out - 1) * stride + 1:stride]  # (B, t_out, Cin)
        y += rows @ w[j]
    if b is not None:
        y += b
    return y


def conv1d_same_backward(x: np.ndarray, w: np.ndarray, dy: np.ndarray, stride: int = 1):
    """Gradients of conv1d_same wrt x, w, bias."""
    B, T, Cin = x.shape
    k, _, Cout = w.shape
    t_out, pad_l, pad_r = same_padding(T, k, stride)
    xp = np.zeros((B, T + pad_l + pad_r, Cin), dtype=x.dtype)
    xp[:, pad_l:pad_l + T] = x
    dxp = np.zeros_like(xp)
    dw = np.zeros_like(w)
    for j in range(k):
        sl = slice(j, j + (t_out - 1) * stride + 1, stride)
        dw[j] = np.tensordot(xp[:, sl], dy, axes=([0, 1], [0, 1]))  # sum_{b,t} x[b,t,i] dy[b,t,o]
        dxp[:, sl] += dy @ w[j].T
    db = dy.sum(axis=(0, 1))
    return dxp[:, pad_l:pad_l + T], dw, db


def conv1d_same_backward_input(w: np.ndarray, dy: np.ndarray, T: int, stride: int = 1) -> np.ndarray:
    """Only the input gradient of conv1d_same (for large shapes where dW is not wanted)."""
    B = dy.shape[0]
    k, Cin, _ = w.shape
    t_out, pad_l, pad_r = same_padding(T, k, stride)
    dxp = np.zeros((B, T + pad_l + pad_r, Cin), dtype=dy.dtype)
    for j in range(k):
        dxp[:, j:j + (t_out - 1) * stride + 1:stride] += dy @ w[j].T
    return dxp[:, pad_l:pad_l + T]


def softmax(z: np.ndarray) -> np.ndarray:
    e = np.exp(z - z.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def glorot_uniform(rng: np.random.Generator, k: int, cin: int, cout: int, dtype=np.float32) -> np.ndarray:
    """Keras default kernel initializer: U(-l, l), l = sqrt(6 / (k*cin + k*cout))."""
    limit = np.sqrt(6.0 / (k * cin + k * cout))
    return rng.uniform(-limit, limit, size=(k, cin, cout)).astype(dtype)


# --------------------------------------------------------------------------------------
# The tower (net.py:291-341)
# --------------------------------------------------------------------------------------
def wav2letter_layer_specs(input_size: int, grapheme_set_size: int, main_filter_count: int = 250,
                           out_filter_count: int = 2000, use_raw_wave_input: bool = False):
    """(name, cin, cout, kernel, stride, activation) for the 11 Conv1D layers of net.py:307-331,
    preceded by `wave_conv` (k250, s160, net.py:310-312) when `use_raw_wave_input`."""
    m, o = main_filter_count, out_filter_count
    specs = [("wave_conv", input_size, m, 250, 160, "relu")] if use_raw_wave_input else []
    specs += [("striding_conv", m if use_raw_wave_input else input_size, m, 48, 2, "relu")]
    specs += [("inner_conv_{}".format(i), m, m, 7, 1, "relu") for i in range(1, 8)]
    specs += [("big_conv_1", m, o, 32, 1, "relu"), ("big_conv_2", o, o, 1, 1, "relu"),
              ("output_conv", o, grapheme_set_size, 1, 1, "softmax")]
    return specs


class Wav2LetterOracle:
    """numpy restatement of the predictive net + CTC objective, with manual backprop."""

    def __init__(self, input_size: int, grapheme_set_size: int, main_filter_count: int = 250,
                 out_filter_count: int = 2000, seed: int = 0, dtype=np.float64,
                 use_raw_wave_input: bool = False):
        self.specs = wav2letter_layer_specs(input_size, grapheme_set_size, main_filter_count, out_filter_count,
                                            use_raw_wave_input)
        self.dtype = dtype
        rng = np.random.default_rng(seed)
        self.weights = [glorot_uniform(rng, k, cin, cout).astype(dtype) for (_, cin, cout, k, _, _) in self.specs]
        self.biases = [np.zeros(cout, dtype=dtype) for (_, _, cout, _, _, _) in self.specs]

    def set_weights(self, weights: Sequence[np.ndarray], biases: Sequence[np.ndarray]):
        self.weights = [np.asarray(w, dtype=self.dtype) for w in weights]
        self.biases = [np.asarray(b, dtype=self.dtype) for b in biases]

    def forward(self, x: np.ndarray, keep: bool = False, dropout_masks=None, dropout_scale: float = 1.0):
        """x (B,T,F) -> probabilities (B, ceil(T/2), V); optionally the per-layer inputs and logits.
        `dropout_masks[i]` (bool, shape of layer i's input) = keep mask of the Dropout layer in front
        of layer i in the training phase (inverted dropout: kept values are scaled by dropout_scale)."""
        a = np.asarray(x, dtype=self.dtype)
        inputs = []
        logits = None
        for index, ((name, _, _, _, stride, act), w, b) in enumerate(zip(self.specs, self.weights, self.biases)):
            if dropout_masks is not None and dropout_masks.get(index) is not None:
                a = a * dropout_masks[index] * dropout_scale
            inputs.append(a)
            z = conv1d_same(a, w, b, stride)
            if act == "relu":
                a = np.maximum(z, 0)
            else:
                logits = z
                a = softmax(z)
        if keep:
            return a, logits, inputs
        return a

    def loss_and_gradients(self, x, labels: np.ndarray, prediction_lengths, label_lengths, dropout_masks=None,
                           dropout_scale: float = 1.0, relu_masks=None):
        """Mean-over-batch CTC objective (net.py:389) and its gradient wrt every kernel / bias.
        `relu_masks` (optional, {layer index: bool (B, T', Cout)}) replaces the oracle's own ReLU sign
        pattern: the gradient is discontinuous where a pre-activation crosses zero, so a reduced-precision
        forward pass that flips a few near-zero units is compared "given the sign pattern it saw"."""
        probs, logits, inputs = self.forward(x, keep=True, dropout_masks=dropout_masks, dropout_scale=dropout_scale)
        B = probs.shape[0]
        losses, dlogits = ctc_batch_cost_with_logit_grad(probs, labels, prediction_lengths, label_lengths)
        d = dlogits / B  # objective = mean_b loss_b
        dws, dbs = [None] * len(self.specs), [None] * len(self.specs)
        for i in reversed(range(len(self.specs))):
            _, _, _, _, stride, act = self.specs[i]
            if act == "relu":
                # d is the gradient wrt this layer's post-activation output
                if relu_masks is not None and relu_masks.get(i) is not None:
                    z_pos = relu_masks[i]
                else:
                    z_pos = conv1d_same(inputs[i], self.weights[i], self.biases[i], stride) > 0
                d = d * z_pos
            dx, dws[i], dbs[i] = conv1d_same_backward(inputs[i], self.weights[i], d, stride)
            if dropout_masks is not None and dropout_masks.get(i) is not None:
                dx = dx * dropout_masks[i] * dropout_scale
            d = dx
        return losses, probs, logits, dws, dbs


# --------------------------------------------------------------------------------------
# CTC  (K.ctc_batch_cost -> tf.nn.ctc_loss; blank = V-1)
# --------------------------------------------------------------------------------------
def _logsumexp(*args):
    m = np.maximum.reduce(args)
    safe = np.where(np.isfinite(m), m, 0.0)
    with np.errstate(divide="ignore"):
        return np.where(np.isfinite(m), safe + np.log(sum(np.exp(a - safe) for a in args)), -np.inf)


def _shift(a: np.ndarray, n: int) -> np.ndarray:
    """out[s] = a[s - n] (n > 0) or a[s + |n|] (n < 0), -inf where out of range."""
    out = np.full_like(a, -np.inf)
    if n > 0:
        out[n:] = a[:len(a) - n] if len(a) > n else []
    else:
        out[:len(a) + n if len(a) + n > 0 else 0] = a[-n:]
    return out


def ctc_log_probs(probs: np.ndarray) -> np.ndarray:
    """What tf.nn.ctc_loss works on after K.ctc_batch_cost: log_softmax(log(p + eps))."""
    u = np.log(probs + EPSILON)
    return u - np.log(np.sum(probs + EPSILON, axis=-1, keepdims=True))


def ctc_alpha_beta(lp: np.ndarray, label: Sequence[int], blank: int):
    """lp (P,V) log-probs of one utterance -> (alpha, beta, log p(label|x)); beta includes the emission at t."""
    P, V = lp.shape
    L = len(label)
    S = 2 * L + 1
    ext = np.full(S, blank, dtype=np.int64)
    ext[1::2] = label
    skip = np.zeros(S, dtype=bool)
    skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
    neg = -np.inf
    alpha = np.full((P, S), neg)
    alpha[0, 0] = lp[0, blank]
    if S > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, P):
        a0 = alpha[t - 1]
        a1 = _shift(a0, 1)
        a2 = np.where(skip, _shift(a0, 2), neg)
        alpha[t] = _logsumexp(a0, a1, a2) + lp[t, ext]
    beta = np.full((P, S), neg)
    beta[P - 1, S - 1] = lp[P - 1, blank]
    if S > 1:
        beta[P - 1, S - 2] = lp[P - 1, ext[S - 2]]
    skip_b = np.zeros(S, dtype=bool)  # may state s jump to s+2 ?
    skip_b[:-2] = skip[2:]
    for t in range(P - 2, -1, -1):
        b0 = beta[t + 1]
        b1 = _shift(b0, -1)
        b2 = np.where(skip_b, _shift(b0, -2), neg)
        beta[t] = _logsumexp(b0, b1, b2) + lp[t, ext]
    ll = alpha[P - 1, S - 1] if S == 1 else _logsumexp(alpha[P - 1, S - 1], alpha[P - 1, S - 2])
    return alpha, beta, float(ll), ext


def ctc_feasible(label: Sequence[int], P: int) -> bool:
    repeats = sum(1 for a, b in zip(label[:-1], label[1:]) if a == b)
    return P >= len(label) + repeats


def ctc_batch_cost_with_logit_grad(probs: np.ndarray, labels: np.ndarray, prediction_lengths, label_lengths,
                                   blank: Optional[int] = None):
    """Per-utterance loss (B,) and d(sum_b loss_b)/d(logits) (B,T',V), chained through
    log(p+eps) and the softmax that produced `probs`."""
    probs = np.asarray(probs)
    B, T, V = probs.shape
    blank = V - 1 if blank is None else blank
    lp_all = ctc_log_probs(probs.astype(np.float64))
    losses = np.zeros(B)
    dlogits = np.zeros((B, T, V))
    for b in range(B):
        P, L = int(prediction_lengths[b]), int(label_lengths[b])
        label = [int(v) for v in labels[b, :L]]
        if not ctc_feasible(label, P):
            raise ValueError("Not enough time for target transition sequence (required: {}, available: {})".format(
                L + sum(1 for a, c in zip(label[:-1], label[1:]) if a == c), P))
        lp = lp_all[b, :P]
        alpha, beta, ll, ext = ctc_alpha_beta(lp, label, blank)
        losses[b] = -ll
        p = probs[b, :P].astype(np.float64)
        with np.errstate(over="ignore", invalid="ignore"):
            contrib = np.exp(alpha + beta - lp[:, ext] - ll)  # (P,S)
        occ = np.zeros((P, V))
        for s, v in enumerate(ext):
            occ[:, v] += contrib[:, s]
        g_u = np.exp(lp) - occ
        dLdp = g_u / (p + EPSILON)
        dlogits[b, :P] = p * (dLdp - np.sum(p * dLdp, axis=-1, keepdims=True))
    return losses, dlogits


def ctc_batch_cost(probs, labels, prediction_lengths, label_lengths) -> np.ndarray:
    return ctc_batch_cost_with_logit_grad(probs, labels, prediction_lengths, label_lengths)[0]


def ctc_brute_force_log_likelihood(lp: np.ndarray, label: Sequence[int], blank: int) -> float:
    """Enumerate all V^P paths (tiny cases only): log sum of path probabilities collapsing to `label`."""
    P, V = lp.shape
    total = -np.inf
    target = list(label)
    for path in product(range(V), repeat=P):
        collapsed = [k for k, _ in groupby(path)]
        collapsed = [c for c in collapsed if c != blank]
        if collapsed == target:
            total = np.logaddexp(total, sum(lp[t, c] for t, c in enumerate(path)))
    return float(total)


# --------------------------------------------------------------------------------------
# Greedy decode (tf.nn.ctc_greedy_decoder(merge_repeated=True) + net.py:436,468-475)
# --------------------------------------------------------------------------------------
def greedy_decode(probs: np.ndarray, prediction_lengths, blank: Optional[int] = None,
                  merge_repeated: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """-> dense int32 (B, T) padded with -1, lengths (B,).  Lowest index wins argmax ties."""
    B, T, V = probs.shape
    blank = V - 1 if blank is None else blank
    out = -np.ones((B, T), dtype=np.int32)
    lens = np.zeros(B, dtype=np.int32)
    for b in range(B):
        prev = -1
        n = 0
        for t in range(int(prediction_lengths[b])):
            c = int(np.argmax(probs[b, t]))
            if c != blank and not (merge_repeated and c == prev):
                out[b, n] = c
                n += 1
            prev = c
        lens[b] = n
    return out, lens


# --------------------------------------------------------------------------------------
# Keras-2 Adam
# --------------------------------------------------------------------------------------
class KerasAdam:
    def __init__(self, lr=1e-4, beta_1=0.9, beta_2=0.999, epsilon=1e-8):
        self.lr, self.beta_1, self.beta_2, self.epsilon = lr, beta_1, beta_2, epsilon
        self.iterations = 0
        self.m = None
        self.v = None

    def step(self, params: List[np.ndarray], grads: List[np.ndarray]) -> List[np.ndarray]:
        if self.m is None:
            self.m = [np.zeros_like(p) for p in params]
            self.v = [np.zeros_like(p) for p in params]
        self.iterations += 1
        t = self.iterations
        lr_t = self.lr * np.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)
        out = []
        for i, (p, g) in enumerate(zip(params, grads)):
            self.m[i] = self.beta_1 * self.m[i] + (1.0 - self.beta_1) * g
            self.v[i] = self.beta_2 * self.v[i] + (1.0 - self.beta_2) * g * g
            out.append(p - lr_t * self.m[i] / (np.sqrt(self.v[i]) + self.epsilon))
        return out


# --------------------------------------------------------------------------------------
# Batching contract (net.py:578-607) and label encoding (grapheme_enconding.py:25-32)
# --------------------------------------------------------------------------------------
def input_batch_and_prediction_lengths(spectrograms: Sequence[np.ndarray], ratio: int = 2):
    lengths = [s.shape[0] for s in spectrograms]
    batch = np.zeros((len(spectrograms), max(lengths), spectrograms[0].shape[1]))
    for i, s in enumerate(spectrograms):
        batch[i, :s.shape[0], :s.shape[1]] = s
    return batch, [n // ratio for n in lengths]


def encode_label_batch(labels: Sequence[str], allowed_characters: Sequence[str]) -> np.ndarray:
    index = {c: i for i, c in enumerate(allowed_characters)}
    out = -np.ones((len(labels), max(len(l) for l in labels)), dtype=np.int32)
    for i, label in enumerate(labels):
        for j, c in enumerate(label):
            if c not in index:
                raise ValueError("Unexpected char: '{}'".format(c))
            out[i, j] = index[c]
    return out


def decode_graphemes(graphemes: Sequence[int], allowed_characters: Sequence[str], merge_repeated: bool = True) -> str:
    """grapheme_enconding.py:34-39,131-137 for the CTC encoding (blank = len(allowed_characters))."""
    blank = len(allowed_characters)
    if merge_repeated:
        graphemes = [k for k, _ in groupby(graphemes)]
    chars = []
    for g in graphemes:
        if 0 <= g < len(allowed_characters):
            chars.append(allowed_characters[g])
        elif g == blank:
            chars.append("")
        else:
            raise ValueError("Unexpected grapheme: '{}'".format(g))
    return "".join(chars)
