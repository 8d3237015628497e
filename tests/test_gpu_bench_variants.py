"""Oracle parity of the exact kernel variants `bench.py --workload full` launches, at the reference
widths (250 / 2000 filters), in every precision mode — through the C-ABI.

The planner (`csrc/api.cu`) only picks these variants on large problems (split-K input gradient needs
>= 148 tiles, K-split weight gradient depends on the K length, tail splitting on the tile count modulo
148), so each one is exercised twice: forced through its tuning variable on a shape the numpy oracle
finishes in seconds, and once on a shape where the planner picks it by itself ("natural").

    split-K dgrad  : conv_gemm_kernel<256, EPI_F32, BMN> + dgrad_finalize_kernel   (big_conv_1)
    wgrad<256>     : K-split TMA reduce-add into dW                                (big_conv_1 / big_conv_2)
    wgrad tap pair : two taps share one N = 256 accumulator                        (striding_conv, 128 mel bins)
    tail split     : narrow tiles in the last partial wave of the persistent grid  (forward / dgrad)
    tail K split   : (opt-in, SL_TAIL_KSPLIT) those tiles split over the contraction instead; partial sums
                     meet in a zero-filled scratch slot and the item with the last ticket folds them in

Tolerances: rel = max|got - want| / max|want| against the fp64 oracle; 1e-4 in the split-bf16 mode
(bf16x2), 2e-2 in bf16 and 3e-3 in fp16 (one rounding of each operand to 8 / 11 mantissa bits).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PRECS = {"bf16": 1, "bf16x2": 2, "fp16": 3}
TOL = {"bf16": 2e-2, "bf16x2": 1e-4, "fp16": 3e-3}


def rel_err(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)


@pytest.fixture(scope="module")
def env():
    import torch
    from oracle import keras_tf_oracle as oracle
    from speechless_b200 import _lib, english_frequent_characters
    from speechless_b200.net import Wav2Letter
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)

    class Env:
        pass

    e = Env()
    e.torch, e.oracle, e.lib, e._lib = torch, oracle, _lib.load(), _lib
    e.alphabet, e.Wav2Letter = english_frequent_characters, Wav2Letter
    return e


def storage(torch, prec):
    return torch.float16 if prec == 3 else torch.bfloat16


def planes(prec):
    return 2 if prec == 2 else 1


def pad64(c):
    return (c + 63) // 64 * 64


def pack(env, a, prec, t_alloc=None):
    """fp32 (B,T,C) numpy -> packed device tensor."""
    torch, lib, check, ptr = env.torch, env.lib, env._lib.check, env._lib.ptr
    B, T, C = a.shape
    t_alloc = t_alloc or T
    d = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to("cuda:0")
    out = torch.zeros((B, t_alloc, planes(prec) * pad64(C)), dtype=storage(torch, prec), device="cuda:0")
    check(lib.sl_pack_activation(ptr(d), ptr(out), B, T, C, t_alloc, pad64(C), prec, None))
    return out


def pack_w(env, w, prec):
    torch, lib, check, ptr = env.torch, env.lib, env._lib.check, env._lib.ptr
    k, cin, cout = w.shape
    d = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)).to("cuda:0")
    out = torch.zeros((k, pad64(cout), planes(prec) * pad64(cin)), dtype=storage(torch, prec), device="cuda:0")
    check(lib.sl_pack_weights(ptr(d), ptr(out), k, cin, cout, pad64(cin), pad64(cout), prec, None))
    check(lib.sl_sync_check())
    return out


def unpack(env, packed, B, T, C, prec, t_alloc=None):
    torch, lib, check, ptr = env.torch, env.lib, env._lib.check, env._lib.ptr
    out = torch.zeros((B, T, C), dtype=torch.float32, device="cuda:0")
    check(lib.sl_unpack_activation(ptr(packed), ptr(out), B, T, C, t_alloc or T, pad64(C), prec, None))
    check(lib.sl_sync_check())
    return out.cpu().numpy()


# ------------------------------------------------------------------ split-K input gradient (big_conv_1)
@pytest.mark.parametrize("mode", ["bf16", "bf16x2", "fp16"])
@pytest.mark.parametrize("B,T,force", [(2, 300, 4), (2, 300, 2), (30, 626, None)])
def test_split_k_input_gradient_big_conv_1(env, monkeypatch, mode, B, T, force):
    """250 -> 2000, k = 32: fp32 partial sums over channel-chunk ranges meet in HBM (TMA reduce-add), then
    dgrad_finalize applies the ReLU mask of the layer below and packs.  (30, 626) is the bench's own
    frame count with 150 tiles >= 148, where the planner splits by itself."""
    if force is None and mode != "bf16":
        pytest.skip("the natural large case runs once, in the benchmark's own mode")
    torch, lib, check, ptr = env.torch, env.lib, env._lib.check, env._lib.ptr
    prec = PRECS[mode]
    cin, cout, k = 250, 2000, 32
    if force is not None:
        monkeypatch.setenv("SL_DGRAD_KSPLIT", str(force))
    need = lib.sl_conv1d_dgrad_workspace_bytes(B, T, cin, cout, k)
    assert need == B * T * 256 * 4, "the planner did not choose the split-K variant for this shape"
    rng = np.random.default_rng(T + B)
    dy = (rng.standard_normal((B, T, cout)) * (rng.random((B, T, cout)) < 0.5)).astype(np.float32)
    w = (rng.standard_normal((k, cin, cout)) / np.sqrt(k * cout)).astype(np.float32)
    below = rng.random((B, T, cin)) < 0.6
    bits = np.zeros((B, T, 256), dtype=np.uint8)
    bits[..., :cin] = below
    mask = torch.from_numpy(np.packbits(bits, axis=2, bitorder="little")).to("cuda:0")
    dyp, wf = pack(env, dy, prec), pack_w(env, w, prec)
    dxp = torch.full((B, T, planes(prec) * 256), 3.0, dtype=storage(torch, prec), device="cuda:0")
    scratch = torch.empty(need, dtype=torch.uint8, device="cuda:0")
    check(lib.sl_conv1d_dgrad(ptr(dyp), ptr(wf), ptr(mask), ptr(dxp), B, T, cin, cout, k, 1, prec, 1.0, ptr(scratch),
                              need, None))
    got = unpack(env, dxp, B, T, cin, prec)
    want = env.oracle.conv1d_same_backward_input(w.astype(np.float64), dy.astype(np.float64), T, 1) * below
    # (tcgen05 accumulates fp32 with truncation: over 32 taps x 2048 channels x 3 terms the bias reaches ~2e-4 of
    # the maximum in the unsplit kernel, DESIGN.md §5, and shrinks with the length of the partial sums)
    assert rel_err(got, want) < max(TOL[mode], 3e-4)
    assert float(dxp.view(B, T, planes(prec), 256)[..., cin:].float().abs().max()) == 0.0  # channel padding stays zero
    # the unsplit kernel (no scratch) gives the same answer up to the rounding of the packed output
    dxp2 = torch.zeros_like(dxp)
    check(lib.sl_conv1d_dgrad(ptr(dyp), ptr(wf), ptr(mask), ptr(dxp2), B, T, cin, cout, k, 1, prec, 1.0, None, 0, None))
    assert rel_err(unpack(env, dxp2, B, T, cin, prec), want) < max(TOL[mode], 5e-4)


# ------------------------------------------------------------------ weight gradients of every bench layer shape
@pytest.mark.parametrize("mode", ["bf16", "bf16x2", "fp16"])
@pytest.mark.parametrize("name,B,T,cin,cout,k,stride,ksplit", [
    ("big_conv_1 K-split 2", 2, 300, 250, 2000, 32, 1, 2),
    ("big_conv_1 K-split 5", 3, 200, 250, 2000, 32, 1, 5),
    ("big_conv_1 planner", 2, 300, 250, 2000, 32, 1, None),
    ("big_conv_2", 2, 333, 2000, 2000, 1, 1, None),
    ("big_conv_2 K-split 3", 2, 333, 2000, 2000, 1, 1, 3),
    ("output_conv", 3, 626, 2000, 29, 1, 1, None),
    ("striding_conv tap pairs, odd T", 2, 1251, 128, 250, 48, 2, None),
    ("striding_conv tap pairs, even T", 3, 300, 128, 250, 48, 2, 2),
    ("inner_conv", 4, 626, 250, 250, 7, 1, None),
])
def test_weight_gradient_bench_shapes(env, monkeypatch, mode, name, B, T, cin, cout, k, stride, ksplit):
    torch, lib, check, ptr = env.torch, env.lib, env._lib.check, env._lib.ptr
    prec = PRECS[mode]
    if ksplit is not None:
        monkeypatch.setenv("SL_WGRAD_KSPLIT", str(ksplit))
    rng = np.random.default_rng(T * 3 + k + B)
    t_out = -(-T // stride)
    t_alloc = (T + stride - 1) // stride * stride
    x = np.maximum(rng.standard_normal((B, T, cin)), 0).astype(np.float32)  # post-ReLU activations
    dy = (rng.standard_normal((B, t_out, cout)) * 0.01).astype(np.float32)
    xp, dyp = pack(env, x, prec, t_alloc), pack(env, dy, prec)
    cip, cop = pad64(cin), pad64(cout)
    dw = torch.full((k, cop, cip), 7.0, dtype=torch.float32, device="cuda:0")
    db = torch.full((cout,), 7.0, dtype=torch.float32, device="cuda:0")
    scale = 0.25  # a power of two, as the fp16 loss scale is
    check(lib.sl_conv1d_wgrad(ptr(xp), ptr(dyp), ptr(dw), ptr(db), B, T, t_alloc, cin, cout, k, stride, prec, 0, scale,
                              None))
    check(lib.sl_sync_check())
    w0 = np.zeros((k, cin, cout))
    _, dw_want, db_want = env.oracle.conv1d_same_backward(x.astype(np.float64), w0, dy.astype(np.float64), stride)
    got = dw.cpu().numpy()  # internal layout (k, cout_pad, cin_pad)
    assert rel_err(got[:, :cout, :cin].transpose(0, 2, 1), dw_want * scale) < TOL[mode], name
    assert rel_err(db.cpu().numpy(), db_want * scale) < TOL[mode], name
    assert np.abs(got[:, cout:, :]).max(initial=0) == 0 and np.abs(got[:, :, cin:]).max(initial=0) == 0
    # accumulate = 1 adds a second copy on top
    check(lib.sl_conv1d_wgrad(ptr(xp), ptr(dyp), ptr(dw), ptr(db), B, T, t_alloc, cin, cout, k, stride, prec, 1, scale,
                              None))
    check(lib.sl_sync_check())
    assert rel_err(dw.cpu().numpy()[:, :cout, :cin].transpose(0, 2, 1), 2 * dw_want * scale) < TOL[mode], name


# ------------------------------------------------------------------ tail of the persistent grid: forward tiles
TAIL_VARIANTS = {"narrow": {}, "k-split": {"SL_TAIL_KSPLIT": "8"}, "whole": {"SL_TAIL_SPLIT": "0"}}


@pytest.mark.parametrize("variant", list(TAIL_VARIANTS))
@pytest.mark.parametrize("mode", ["bf16", "bf16x2", "fp16"])
@pytest.mark.parametrize("B,T,cin,cout,k", [
    (20, 1000, 250, 250, 7),   # inner_conv: 160 tiles = 148 + 12 -> the 12 are split over their 4 channel chunks
    (3, 900, 250, 2000, 32),   # big_conv_1 (A-halo loop): 24 x 8 = 192 tiles = 148 + 44 -> split in 2
    (3, 900, 2000, 2000, 1),   # big_conv_2: 192 tiles, 32 channel chunks -> split in 3 (11 + 11 + 10 chunks)
])
def test_forward_tail_split_tiles(env, monkeypatch, variant, mode, B, T, cin, cout, k):
    torch, lib, check, ptr = env.torch, env.lib, env._lib.check, env._lib.ptr
    if variant == "whole" and (mode != "fp16" or cin == 2000):
        pytest.skip("the unsplit variant runs once per shape, in the benchmark's mode")
    for key, value in TAIL_VARIANTS[variant].items():
        monkeypatch.setenv(key, value)
    prec = PRECS[mode]
    tiles = B * -(-T // 128) * (pad64(cout) // min(pad64(cout), 256))
    assert tiles > 148 and tiles % 148 != 0, "shape must leave a partial last wave"
    rng = np.random.default_rng(T + k)
    x = rng.standard_normal((B, T, cin)).astype(np.float32)
    w = (rng.standard_normal((k, cin, cout)) / np.sqrt(k * cin)).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    xp, wf = pack(env, x, prec), pack_w(env, w, prec)
    bd = torch.from_numpy(bias).to("cuda:0")
    cop = pad64(cout)
    want = np.maximum(env.oracle.conv1d_same(x.astype(np.float64), w.astype(np.float64), bias.astype(np.float64), 1), 0)
    first = None
    for launch in range(3):  # the scratch slots must be all zeros again after every launch
        yp = torch.zeros((B, T, planes(prec) * cop), dtype=storage(torch, prec), device="cuda:0")
        mask = torch.zeros((B, T, cop // 8), dtype=torch.uint8, device="cuda:0")
        check(lib.sl_conv1d_fwd(ptr(xp), ptr(wf), ptr(bd), ptr(yp), ptr(mask), None, None, None, B, T, T, T, cin, cout, k,
                                1, 1, prec, None))
        got = unpack(env, yp, B, T, cout, prec)
        assert rel_err(got, want) < TOL[mode]
        bits = np.unpackbits(mask.cpu().numpy(), axis=2, bitorder="little")[..., :cout].astype(bool)
        assert (bits == (got > 0))[np.abs(want) > 1e-3].all()
        if first is None:
            first = got
        else:
            # the order in which partial sums arrive may change the fp32 rounding of a split tile, i.e. at most
            # the last bit of a packed output value
            assert rel_err(got, first) < 0.5 * TOL[mode]


@pytest.mark.parametrize("mode", ["bf16", "bf16x2", "fp16"])
@pytest.mark.parametrize("B,T,cin,cout,k", [
    (20, 1000, 250, 250, 7),   # inner_conv input gradient: 160 tiles
    (3, 900, 2000, 2000, 1),   # big_conv_2 input gradient: 192 tiles
])
def test_input_gradient_tail_k_split(env, monkeypatch, mode, B, T, cin, cout, k):
    torch, lib, check, ptr = env.torch, env.lib, env._lib.check, env._lib.ptr
    prec = PRECS[mode]
    rng = np.random.default_rng(T + k + 1)
    dy = (rng.standard_normal((B, T, cout)) * (rng.random((B, T, cout)) < 0.5)).astype(np.float32)
    w = (rng.standard_normal((k, cin, cout)) / np.sqrt(k * cout)).astype(np.float32)
    below = rng.random((B, T, cin)) < 0.6
    cip = pad64(cin)
    bits = np.zeros((B, T, cip), dtype=np.uint8)
    bits[..., :cin] = below
    mask = torch.from_numpy(np.packbits(bits, axis=2, bitorder="little")).to("cuda:0")
    dyp, wf = pack(env, dy, prec), pack_w(env, w, prec)
    want = env.oracle.conv1d_same_backward_input(w.astype(np.float64), dy.astype(np.float64), T, 1) * below
    results = {}
    for variant in ("k-split", "narrow"):
        if variant == "k-split":
            monkeypatch.setenv("SL_TAIL_KSPLIT", "8")
        else:
            monkeypatch.delenv("SL_TAIL_KSPLIT")
        for launch in range(2):
            dxp = torch.full((B, T, planes(prec) * cip), 3.0, dtype=storage(torch, prec), device="cuda:0")
            check(lib.sl_conv1d_dgrad(ptr(dyp), ptr(wf), ptr(mask), ptr(dxp), B, T, cin, cout, k, 1, prec, 1.0, None, 0,
                                      None))
            got = unpack(env, dxp, B, T, cin, prec)
            assert rel_err(got, want) < max(TOL[mode], 3e-4), (variant, launch)
            assert float(dxp.view(B, T, planes(prec), cip)[..., cin:].float().abs().max()) == 0.0
        results[variant] = got
    assert rel_err(results["k-split"], results["narrow"]) < max(TOL[mode], 3e-4)


# ------------------------------------------------------------------ narrowed persistent grids (data-parallel window)
@pytest.mark.parametrize("B,T,cin,cout,k", [
    (20, 1000, 250, 250, 7),   # inner_conv (A-halo loop): 160 tiles
    (3, 900, 2000, 2000, 1),   # big_conv_2: 192 tiles
])
def test_forward_on_a_narrowed_grid_is_bitwise_the_same(env, B, T, cin, cout, k):
    """`sl_set_sm_limit` (the window in which a gradient bucket's all-reduce shares the GPU) only
    changes which CTA computes which tile — and which tiles of the last wave are cut narrower — never a value."""
    torch, lib, check, ptr = env.torch, env.lib, env._lib.check, env._lib.ptr
    prec = PRECS["fp16"]
    rng = np.random.default_rng(T + k + 5)
    x = rng.standard_normal((B, T, cin)).astype(np.float32)
    w = (rng.standard_normal((k, cin, cout)) / np.sqrt(k * cin)).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    xp, wf = pack(env, x, prec), pack_w(env, w, prec)
    bd = torch.from_numpy(bias).to("cuda:0")
    cop = pad64(cout)
    outputs = []
    try:
        for limit in (0, 132, 100, 37):
            check(lib.sl_set_sm_limit(limit))
            yp = torch.zeros((B, T, cop), dtype=storage(torch, prec), device="cuda:0")
            mask = torch.zeros((B, T, cop // 8), dtype=torch.uint8, device="cuda:0")
            check(lib.sl_conv1d_fwd(ptr(xp), ptr(wf), ptr(bd), ptr(yp), ptr(mask), None, None, None, B, T, T, T, cin, cout,
                                    k, 1, 1, prec, None))
            check(lib.sl_sync_check())
            outputs.append((yp.clone(), mask.clone()))
    finally:
        check(lib.sl_set_sm_limit(0))
    for yp, mask in outputs[1:]:
        assert torch.equal(yp, outputs[0][0]) and torch.equal(mask, outputs[0][1])
    want = np.maximum(env.oracle.conv1d_same(x.astype(np.float64), w.astype(np.float64), bias.astype(np.float64), 1), 0)
    assert rel_err(unpack(env, outputs[0][0], B, T, cout, prec), want) < TOL["fp16"]


# ------------------------------------------------------------------ full-width tower: logits and gradients per mode
def _tower_case(env, mode, seed=1):
    from tests.test_gpu_parity import make_pair
    from speechless_b200.synthetic import synthetic_batch
    net, ref = make_pair(env, main=250, out=2000, seed=seed, dtype=mode)
    batch = synthetic_batch(2, [203, 171], env.alphabet, seed=6, label_length=14)
    inputs, _ = net._inputs_for_loss_net(batch)
    return net, ref, inputs


def _device_gradients(env, net, inputs):
    names = env.Wav2Letter.InputNames
    tower = net.tower
    ws = tower.upload(inputs[names.input_batch])
    tower.forward(ws, want_logits=True)
    tower.set_labels(ws, inputs[names.label_batch], inputs[names.prediction_lengths], inputs[names.label_lengths])
    loss = tower.ctc(ws, want_grad=True, grad_scale=1.0 / ws.B)
    tower.backward(ws)
    tower.sync()
    saved = tower.params.clone()
    tower.params.copy_(tower.grads)  # read the gradients back through the Keras-layout accessor
    grads = [tower.get_layer_weights(i) for i in range(len(tower.layers))]
    tower.params.copy_(saved)
    masks = {}
    for index, layer in enumerate(tower.layers[:-1]):
        bits = np.unpackbits(ws.masks[index].cpu().numpy(), axis=2, bitorder="little")
        masks[index] = bits[..., :layer.cout].astype(bool)
    return ws.logits.cpu().numpy(), loss.cpu().numpy(), grads, masks


@pytest.mark.parametrize("mode,logit_tol,grad_tol,grad_tol_given_masks", [
    ("bf16x2", 1e-3, None, 5e-3),   # fp32-parity mode: BASELINE.json's tolerance
    ("fp16", 1e-3, None, 5e-3),     # logits meet 1e-3 at the bf16 cost
    ("bf16", 5e-2, None, 5e-2),     # throughput mode of BASELINE config 3: reported, bounded loosely
])
def test_full_width_tower_logits_and_gradients(env, mode, logit_tol, grad_tol, grad_tol_given_masks):
    """Reference widths 250 / 2000, V = 29, two ragged utterances: logits, per-utterance loss and the gradient
    of every kernel and bias against the fp64 oracle.  A reduced-precision forward pass flips the sign of a
    few near-zero pre-activations; the gradient is discontinuous there (a whole unit switches on or off: measured
    on B200, even the split-bf16 mode with its ~1e-5 activation error flips a handful of the 1.5 M ReLU signs of
    this case, which moves the worst bias-gradient entry by 7 %), so every mode is compared with the oracle's
    gradient GIVEN the ReLU sign pattern the device saw, and the raw error (flips included) is printed."""
    net, ref, inputs = _tower_case(env, mode)
    names = env.Wav2Letter.InputNames
    logits, loss, grads, masks = _device_gradients(env, net, inputs)
    x = inputs[names.input_batch].astype(np.float64)
    args = (inputs[names.label_batch], inputs[names.prediction_lengths][:, 0], inputs[names.label_lengths][:, 0])
    losses_ref, _, logits_ref, dws, dbs = ref.loss_and_gradients(x, *args)
    _, _, _, dws_m, dbs_m = ref.loss_and_gradients(x, *args, relu_masks=masks)
    logit_err = rel_err(logits, logits_ref)
    raw = max(max(rel_err(g[0], dws[i]), rel_err(g[1], dbs[i])) for i, g in enumerate(grads))
    given = max(max(rel_err(g[0], dws_m[i]), rel_err(g[1], dbs_m[i])) for i, g in enumerate(grads))
    flips = sum(int((masks[i] != (ref_mask > 0)).sum()) for i, ref_mask in
                enumerate(_oracle_pre_activations(env, ref, x)) if i in masks)
    total = sum(m.size for m in masks.values())
    print("{}: logits rel err {:.2e}, loss rel err {:.2e}, gradient rel err raw {:.2e} / given the device's ReLU "
          "pattern {:.2e}, {} of {} ReLU signs differ".format(mode, logit_err, np.abs(loss / losses_ref - 1).max(), raw,
                                                              given, flips, total))
    assert logit_err < logit_tol
    assert np.abs(loss / losses_ref - 1).max() < (1e-4 if mode != "bf16" else 1e-3)
    assert given < grad_tol_given_masks
    if grad_tol is not None:
        assert raw < grad_tol


def _oracle_pre_activations(env, ref, x):
    _, _, layer_inputs = ref.forward(x, keep=True)
    return [env.oracle.conv1d_same(layer_inputs[i], ref.weights[i], ref.biases[i], ref.specs[i][4])
            for i in range(len(ref.specs))]


def test_fp16_loss_scale_is_exact_and_protects_small_gradients(env):
    """The fp16 mode multiplies dlogits by a power of two S and the weight-gradient epilogue by 1/S: the result
    does not depend on S while nothing under/overflows, and with a global batch of 4096 utterances (dlogits
    <= 2.4e-4, deep inside fp16's subnormals after a few layers) the scaled path still matches the oracle."""
    net, ref, inputs = _tower_case(env, "fp16", seed=3)
    names = env.Wav2Letter.InputNames
    tower = net.tower
    x = inputs[names.input_batch].astype(np.float64)
    args = (inputs[names.label_batch], inputs[names.prediction_lengths][:, 0], inputs[names.label_lengths][:, 0])

    def grads_with(target, grad_scale):
        tower.loss_scale_target = target
        ws = tower.upload(inputs[names.input_batch])
        tower.forward(ws)
        tower.set_labels(ws, inputs[names.label_batch], inputs[names.prediction_lengths], inputs[names.label_lengths])
        tower.ctc(ws, want_grad=True, grad_scale=grad_scale)
        tower.backward(ws)
        tower.sync()
        masks = {i: np.unpackbits(ws.masks[i].cpu().numpy(), axis=2, bitorder="little")[..., :l.cout].astype(bool)
                 for i, l in enumerate(tower.layers[:-1])}
        return tower.grads.clone(), ws.loss_scale, masks

    g256, s256, masks = grads_with(256.0, 0.5)
    g16, s16, _ = grads_with(16.0, 0.5)
    assert s256 == 512.0 and s16 == 32.0
    scale = float(g256.abs().max())
    assert float((g256 - g16).abs().max()) < 2e-3 * scale  # same gradient, different subnormal losses only
    # tiny per-utterance weight (as in a 4096-utterance global batch)
    tiny = 1.0 / 4096
    g_scaled, s, _ = grads_with(256.0, tiny)
    assert s == 2.0 ** 20
    g_unscaled, s1, _ = grads_with(0.0, tiny)
    assert s1 == 1.0
    _, _, _, dws, dbs = ref.loss_and_gradients(x, *args, relu_masks=masks)
    first = tower.layers[0]
    want = dws[0] * (2 * tiny)  # oracle objective = mean over B = 2 -> rescale to grad_scale = tiny
    view = lambda g: g[first.w_offset:first.w_offset + first.w_size].view(first.kernel, first.cout_pad,
                                                                          first.cin_pad)[:, :first.cout, :first.cin] \
        .permute(0, 2, 1).cpu().numpy()
    err_scaled, err_unscaled = rel_err(view(g_scaled), want), rel_err(view(g_unscaled), want)
    print("striding_conv dW rel err at grad_scale 1/4096: loss-scaled {:.2e}, unscaled {:.2e}".format(
        err_scaled, err_unscaled))
    assert err_scaled < 5e-3
    assert err_unscaled > err_scaled  # what the scale is for
