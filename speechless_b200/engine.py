"""Device-side engine of the wav2letter hot path: owns the HBM layout and drives the C-ABI.

This is what replaces the Keras `Sequential` of reference net.py:291-341 and the TF graph
behind `loss_net.fit_generator` / `backend.function` (net.py:350-390,456-459,550).  PyTorch
is used for storage (`torch.empty`), streams and `torch.distributed` only; every arithmetic
operation on the path is a kernel of libspeechless_b200.so.

HBM layout
  params / grads / adam_m / adam_v : one flat fp32 buffer each; per layer the kernel in the
        *internal master layout* (k, cout_pad, cin_pad) followed by the bias (cout_pad).
        Channel pads are zero and stay zero (their gradients are exactly zero).
  w_fwd[l]  (k, cout_pad, planes*cin_pad)  bf16   B operand of the forward implicit GEMM (K-major)
        and of the input-gradient GEMM (the same bytes read MN-major) — refreshed by the fused Adam kernel
  act[l]    (B, T', planes*cout_pad)       bf16   post-ReLU output of layer l (kept for backward)
  probs (B,T',V) fp32, logp (B,T',64) fp32, CTC lattices alpha/beta (B,T',S_pad) fp32.
"""
import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from speechless_b200 import _lib
from speechless_b200._lib import ACT_NONE, ACT_RELU, ACT_SOFTMAX, PREC_BF16, PREC_BF16X2, PREC_FP16, check, ptr


def round_up(value: int, multiple: int) -> int:
    return (value + multiple - 1) // multiple * multiple


@dataclass
class LayerSpec:
    name: str
    cin: int
    cout: int
    kernel: int
    stride: int
    activation: str  # "relu" | "linear" | "softmax"
    w_offset: int = 0  # offsets into the flat parameter buffer (floats)
    b_offset: int = 0
    # `windowed`: the layer's receptive fields are laid out as rows first (sl_window_activation), so
    # the GEMM sees one tap over kernel*cin channels at stride 1 (raw-wave `wave_conv`, net.py:310-312)
    windowed: bool = False

    @property
    def gemm_cin(self) -> int:
        return self.kernel * self.cin if self.windowed else self.cin

    @property
    def gemm_kernel(self) -> int:
        return 1 if self.windowed else self.kernel

    @property
    def gemm_stride(self) -> int:
        return 1 if self.windowed else self.stride

    @property
    def cin_pad(self) -> int:
        return round_up(self.gemm_cin, 64)

    @property
    def cout_pad(self) -> int:
        return round_up(self.cout, 64)

    @property
    def w_size(self) -> int:
        return self.gemm_kernel * self.cout_pad * self.cin_pad


def wav2letter_layers(input_size_per_time_step: int, grapheme_set_size: int, activation: str = "relu",
                      output_activation: str = "softmax", main_filter_count: int = 250,
                      out_filter_count: int = 2000, use_raw_wave_input: bool = False) -> List[LayerSpec]:
    """The 11 Conv1D layers of `Wav2Letter.create_predictive_net` (reference net.py:307-331):
    striding_conv k48 s2, inner_conv_1..7 k7, big_conv_1 k32, big_conv_2 k1, output_conv k1;
    with `use_raw_wave_input` a twelfth, `wave_conv` k250 s160, in front (net.py:310-312)."""
    m, o = main_filter_count, out_filter_count
    layers = []
    if use_raw_wave_input:
        layers.append(LayerSpec("wave_conv", input_size_per_time_step, m, 250, 160, activation, windowed=True))
    layers.append(LayerSpec("striding_conv", m if use_raw_wave_input else input_size_per_time_step, m, 48, 2,
                            activation))
    layers += [LayerSpec("inner_conv_{}".format(i), m, m, 7, 1, activation) for i in range(1, 8)]
    layers += [LayerSpec("big_conv_1", m, o, 32, 1, activation),
               LayerSpec("big_conv_2", o, o, 1, 1, activation),
               LayerSpec("output_conv", o, grapheme_set_size, 1, 1, output_activation)]
    offset = 0
    for layer in layers:
        layer.w_offset = offset
        offset += layer.w_size
        layer.b_offset = offset
        offset += layer.cout_pad
    return layers


def same_padding(T: int, k: int, stride: int) -> Tuple[int, int]:
    """TF SAME rule -> (T_out, pad_left)."""
    t_out = -(-T // stride)
    total = max((t_out - 1) * stride + k - T, 0)
    return t_out, total // 2


def ctc_required_frames(label: Sequence[int]) -> int:
    """tf.nn.ctc_loss needs len(label) + #adjacent repeats frames."""
    return len(label) + sum(1 for a, b in zip(label[:-1], label[1:]) if a == b)


class _Arena:
    """Named flat buffers of one tower that only grow.  Real corpora give almost every batch its own
    longest utterance, i.e. its own (B, T): the per-shape workspaces below are therefore only VIEWS of these
    buffers (sized by the largest shape seen so far), so a new shape costs a few Python objects instead of a
    gigabyte of cudaMalloc + cudaHostAlloc per training step.  `generation` changes whenever a buffer had to
    be replaced by a larger one; views taken before that are stale and get rebuilt."""

    def __init__(self, device: torch.device):
        self.device = device
        self.buffers: Dict[str, torch.Tensor] = {}
        self.generation = 0

    def take(self, name: str, shape: Sequence[int], dtype: torch.dtype, pinned: bool = False) -> torch.Tensor:
        count = 1
        for extent in shape:
            count *= int(extent)
        nbytes = max(count * torch.empty((), dtype=dtype).element_size(), 16)
        buffer = self.buffers.get(name)
        if buffer is None or buffer.numel() < nbytes:
            size = nbytes + nbytes // 8  # headroom: the next slightly longer batch does not reallocate again
            buffer = torch.empty(size, dtype=torch.uint8, pin_memory=True) if pinned else \
                torch.empty(size, dtype=torch.uint8, device=self.device)
            self.buffers[name] = buffer
            self.generation += 1
        return buffer[:count * torch.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)


class _Workspace:
    """Views of the tower's arena for one (B, T) batch shape (the small per-utterance vectors are its own).
    Only one shape is in use on the compute stream at a time; the two input slots are separate buffers, so
    the host->device copy of the NEXT batch (any shape) overlaps the step of the current one."""

    def __init__(self, tower: "ConvTower", B: int, T: int):
        dev, planes = tower.device, tower.planes
        self._views: List[Tuple[str, torch.Tensor]] = []
        self._arena = tower.arena
        arena = self  # (take() below records which arena buffers this workspace looks at)
        first = tower.layers[0]
        self.B, self.T = B, T
        # T0: rows per utterance of the packed first-layer operand (frames, or receptive-field rows
        # of a windowed raw-wave layer)
        self.T0 = same_padding(T, first.kernel, first.stride)[0] if first.windowed else T
        self.T_alloc = round_up(self.T0, first.gemm_stride)
        # two input slots for the pipelined training loop (copy of batch i+1 overlaps step i); slot 0 doubles
        # as the plain upload() target
        self.x_slots = [arena.take("x_slot{}".format(i), (B, T, first.cin), torch.float32) for i in range(2)]
        self.host_slots = [arena.take("host_slot{}".format(i), (B, T, first.cin), torch.float32, pinned=True)
                           for i in range(2)]
        self.x_f32, self.x_host = self.x_slots[0], self.host_slots[0]
        # (the pack kernel writes every row of x_packed, allocation padding included)
        self.x_packed = arena.take("x_packed", (B, self.T_alloc, planes * first.cin_pad), tower.storage_dtype)
        self.t_out: List[int] = []
        self.acts: List[torch.Tensor] = []
        t = T
        for layer in tower.layers:
            t, _ = same_padding(t, layer.kernel, layer.stride)
            self.t_out.append(t)
        self.Tp = self.t_out[-1]
        self.masks: List[torch.Tensor] = []  # ReLU sign bits of act[l], 1 bit per (frame, channel)
        # rows allocated per utterance for act[l]: a stride-2 consumer reads it through an
        # (even, odd) frame view, so an odd frame count gets one extra row that stays zero
        self.t_alloc_out: List[int] = []
        for index, (layer, t_out) in enumerate(zip(tower.layers[:-1], self.t_out[:-1])):
            alloc = round_up(t_out, tower.layers[index + 1].gemm_stride)
            self.t_alloc_out.append(alloc)
            self.acts.append(arena.take("act{}".format(index), (B, alloc, planes * layer.cout_pad), tower.storage_dtype))
            self.masks.append(arena.take("mask{}".format(index), (B, t_out, layer.cout_pad // 8), torch.uint8))
        V = tower.layers[-1].cout
        self.probs = arena.take("probs", (B, self.Tp, V), torch.float32)
        self.logits = arena.take("logits", (B, self.Tp, V), torch.float32)
        self.logp = arena.take("logp", (B, self.Tp, 64), torch.float32)
        self.loss = torch.empty((B,), dtype=torch.float32, device=dev)
        self.input_len = torch.empty((B,), dtype=torch.int32, device=dev)
        self.label_len = torch.empty((B,), dtype=torch.int32, device=dev)
        self.decoded = arena.take("decoded", (B, self.Tp), torch.int32)
        self.decoded_len = torch.empty((B,), dtype=torch.int32, device=dev)
        self.labels: Optional[torch.Tensor] = None
        self.ctc_ws: Optional[torch.Tensor] = None
        self.dz_packed: Optional[torch.Tensor] = None
        self.dz_f32: Optional[torch.Tensor] = None
        self.dact: List[Optional[torch.Tensor]] = [None, None]
        self.dgrad_ws: Optional[torch.Tensor] = None  # fp32 scratch of the split-K input gradient
        # training-phase dropout: dropped copy of a layer's input and the combined (keep & ReLU) mask
        self.xdrop: Dict[int, torch.Tensor] = {}
        self.bwd_mask: Dict[int, torch.Tensor] = {}
        self.layer_inputs: List[Optional[torch.Tensor]] = [None] * len(tower.layers)
        self.dropout_seeds: Dict[int, int] = {}
        self.input_dropped = False  # raw-wave input: dropout was applied while packing
        self.loss_scale = 1.0  # power of two the packed gradients of the last ctc() call carry (fp16 mode)

    def take(self, name: str, shape: Sequence[int], dtype: torch.dtype, pinned: bool = False) -> torch.Tensor:
        view = self._arena.take(name, shape, dtype, pinned)
        self._views.append((name, view))
        return view

    def stale(self) -> bool:
        """True once the arena has replaced (grown) one of the buffers this workspace holds views of."""
        buffers = self._arena.buffers
        return any(buffers[name].data_ptr() != view.data_ptr() for name, view in self._views)

    def zero_allocation_rows(self) -> None:
        """The extra row of an activation whose consumer strides by 2 over an odd frame count must be zero, and
        no kernel writes it: zero it when this shape takes over buffers another shape has used."""
        for act, t_out, alloc in zip(self.acts, self.t_out, self.t_alloc_out):
            if alloc > t_out:
                act[:, t_out:].zero_()

    def ensure_backward(self, tower: "ConvTower"):
        if self.dz_packed is None:
            planes = tower.planes
            self.dz_packed = self.take("dz_packed", (self.B, self.Tp, planes * 64), tower.storage_dtype)
            widest = max(layer.cout_pad for layer in tower.layers[:-1])
            rows = max(self.t_out)
            for i in range(2):
                self.dact[i] = self.take("dact{}".format(i), (self.B, rows, planes * widest), tower.storage_dtype)



class ConvTower:
    """Weights + kernels of the 11-layer tower, CTC objective and Keras-2 Adam on one GPU."""

    def __init__(self, layers: List[LayerSpec], device: torch.device, precision: int = PREC_BF16X2,
                 frozen_layer_count: int = 0, dropout: Optional[float] = None,
                 dropout_layers: Sequence[int] = (), dropout_seed: int = 0,
                 loss_scale_target: Optional[float] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("speechless_b200 needs a CUDA device (B200); there is no CPU fallback.")
        self.lib = _lib.load()
        self.layers = layers
        self.device = device
        self.precision = precision
        self.planes = 2 if precision == PREC_BF16X2 else 1
        # fp16 operands: 11 mantissa bits at the bf16 cost, but 5 exponent bits.  Activations of a
        # z-normalised input stay far inside the range; the back-propagated gradients (<= 1/B at the
        # logits) would sink into the subnormals, so they carry a power-of-two LOSS SCALE: the CTC
        # kernel emits dlogits * S with S chosen so that |dlogits * S| <= loss_scale_target, every
        # input gradient inherits the factor, and the weight-gradient epilogue multiplies by 1/S
        # (exact).  bf16 modes have the fp32 range and run with S = 1.
        self.storage_dtype = torch.float16 if precision == PREC_FP16 else torch.bfloat16
        if loss_scale_target is None:
            loss_scale_target = float(os.environ.get("SL_LOSS_SCALE_TARGET", "256")) if precision == PREC_FP16 else 0.0
        self.loss_scale_target = loss_scale_target
        self.frozen_layer_count = frozen_layer_count
        # inverted dropout in front of the listed conv layers, training phase only (net.py:301-303);
        # the kernel quantises p to 16 bits, the scale below matches it exactly
        self.dropout = dropout if dropout else None
        self.dropout_layers = set(dropout_layers) if self.dropout else set()
        self.dropout_seed = dropout_seed
        self.dropout_step = 0
        if self.dropout is not None and not (0.0 < self.dropout < 1.0):
            raise ValueError("dropout rate must be in (0, 1)")
        self.dropout_scale = 1.0 / (1.0 - int(self.dropout * 65536.0 + 0.5) / 65536.0) if self.dropout else 1.0
        for layer in layers:
            if layer.activation not in ("relu", "linear", "softmax"):
                raise NotImplementedError("activation '{}' has no sm_100a epilogue".format(layer.activation))
            if layer.gemm_stride not in (1, 2):
                raise NotImplementedError("stride {} needs a windowed layer".format(layer.stride))
        if any(layer.windowed for layer in layers[1:]):
            raise NotImplementedError("only the first layer can be windowed")
        if layers[-1].activation != "softmax":
            raise NotImplementedError("the output layer must be a softmax (CTC objective)")
        self.param_count = layers[-1].b_offset + layers[-1].cout_pad
        with torch.cuda.device(device):
            self.params = torch.zeros(self.param_count, dtype=torch.float32, device=device)
            self.grads: Optional[torch.Tensor] = None
            self.adam_m: Optional[torch.Tensor] = None
            self.adam_v: Optional[torch.Tensor] = None
            self.w_fwd = [torch.zeros((l.gemm_kernel, l.cout_pad, self.planes * l.cin_pad), dtype=self.storage_dtype,
                                      device=device) for l in layers]
        self.arena = _Arena(device)
        self._workspaces: Dict[Tuple[int, int], _Workspace] = {}
        self._bound: Optional[_Workspace] = None  # the workspace whose shape the arena buffers currently carry
        self._slot_copied = [None, None]    # event on the copy stream: input slot holds the new batch
        self._slot_consumed = [None, None]  # event on the compute stream: the pack kernel has read the slot
        self._current: Optional[_Workspace] = None
        self.launches = 0  # kernels launched through the C-ABI (bench.py reports it)
        # optional per-kernel timing: list of (kind, layer name, start event, stop event) on the
        # launching stream; bench.py turns it on to measure roofline fractions live
        self.profile: Optional[list] = None
        self.nvtx = os.environ.get("SL_NVTX", "0") == "1"
        self.overlap_backward = False
        self._adam_tables = None
        self._copy_stream = None
        self._side_stream = None
        self._update_stream = None
        self._sm_limit_until = 0  # launch count at which sl_set_sm_limit is lifted again
        # The weight-gradient kernels accumulate into `grads` (TMA reduce-add), so the buffer must be zero when
        # backward starts.  `backward_and_update` zeroes each bucket on the update stream right after its Adam
        # step — off the critical path — unless `retain_gradients` asks for the gradients to stay readable.
        self.retain_gradients = False
        self._grads_clean = False
        # bucketed all-reduce + Adam on an update stream, overlapped with backward (backward_and_update)
        self.pipeline_update = os.environ.get("SL_PIPELINE_UPDATE", "1") != "0"
        self.sm_count = torch.cuda.get_device_properties(device).multi_processor_count

    # ------------------------------------------------------------------ helpers
    @property
    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _timed(self, kind: str, name: str, rc_fn) -> None:
        """Run one C-ABI launch (rc_fn returns its code), bracketed by CUDA events when profiling and by an
        NVTX range (`kind:layer`) when SL_NVTX=1, so that a timeline shows the step's structure."""
        if self.nvtx:
            torch.cuda.nvtx.range_push("{}:{}".format(kind, name))
            try:
                check(rc_fn())
            finally:
                torch.cuda.nvtx.range_pop()
            return
        if self.profile is None:
            check(rc_fn())
            return
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream = torch.cuda.current_stream(self.device)
        start.record(stream)
        check(rc_fn())
        stop.record(stream)
        self.profile.append((kind, name, start, stop))

    def _w(self, buf: torch.Tensor, layer: LayerSpec) -> torch.Tensor:
        return buf[layer.w_offset:layer.w_offset + layer.w_size]

    def _b(self, buf: torch.Tensor, layer: LayerSpec) -> torch.Tensor:
        return buf[layer.b_offset:layer.b_offset + layer.cout_pad]

    def first_trainable(self) -> int:
        return min(self.frozen_layer_count, len(self.layers))

    # ------------------------------------------------------------------ weights
    def init_glorot(self, seed: Optional[int] = None) -> None:
        """Keras defaults: glorot_uniform kernels, zero biases."""
        rng = np.random.default_rng(seed)
        for index, layer in enumerate(self.layers):
            limit = np.sqrt(6.0 / (layer.kernel * layer.cin + layer.kernel * layer.cout))
            kernel = rng.uniform(-limit, limit, size=(layer.kernel, layer.cin, layer.cout)).astype(np.float32)
            self.set_layer_weights(index, kernel, np.zeros(layer.cout, dtype=np.float32), repack=False)
        self.repack()

    def set_layer_weights(self, index: int, kernel: np.ndarray, bias: np.ndarray, repack: bool = True) -> None:
        """kernel in the Keras layout (k, Cin, Cout), bias (Cout,)."""
        layer = self.layers[index]
        kernel = np.ascontiguousarray(kernel, dtype=np.float32)
        bias = np.ascontiguousarray(bias, dtype=np.float32)
        if kernel.shape != (layer.kernel, layer.cin, layer.cout) or bias.shape != (layer.cout,):
            raise ValueError("layer {} expects kernel {} and bias {}, got {} and {}".format(
                layer.name, (layer.kernel, layer.cin, layer.cout), (layer.cout,), kernel.shape, bias.shape))
        with torch.cuda.device(self.device):
            w_keras = torch.from_numpy(kernel).to(self.device)
            # (k, Cin, Cout) row-major is also (1, k*Cin, Cout): a windowed layer needs no reshuffle
            check(self.lib.sl_weights_keras_to_internal(ptr(w_keras), ptr(self._w(self.params, layer)),
                                                        layer.gemm_kernel, layer.gemm_cin, layer.cout, layer.cin_pad,
                                                        layer.cout_pad, self.stream))
            b = self._b(self.params, layer)
            b.zero_()
            b[:layer.cout].copy_(torch.from_numpy(bias).to(self.device))
            if repack:
                self.repack([index])
            torch.cuda.current_stream(self.device).synchronize()  # w_keras must outlive the kernel

    def get_layer_weights(self, index: int) -> List[np.ndarray]:
        layer = self.layers[index]
        with torch.cuda.device(self.device):
            w_keras = torch.empty((layer.kernel, layer.cin, layer.cout), dtype=torch.float32, device=self.device)
            check(self.lib.sl_weights_internal_to_keras(ptr(self._w(self.params, layer)), ptr(w_keras),
                                                        layer.gemm_kernel, layer.gemm_cin, layer.cout, layer.cin_pad,
                                                        layer.cout_pad, self.stream))
            return [w_keras.cpu().numpy(), self._b(self.params, layer)[:layer.cout].cpu().numpy()]

    def repack(self, indices: Optional[Sequence[int]] = None) -> None:
        """fp32 master -> bf16 tensor-core operands (after a weight load; the optimizer step
        refreshes them itself)."""
        with torch.cuda.device(self.device):
            for index in (range(len(self.layers)) if indices is None else indices):
                layer = self.layers[index]
                check(self.lib.sl_pack_weights_internal(ptr(self._w(self.params, layer)), ptr(self.w_fwd[index]),
                                                        layer.gemm_kernel, layer.cin_pad, layer.cout_pad,
                                                        self.precision, self.stream))
                self.launches += 1

    def broadcast_parameters(self, src: int = 0) -> None:
        """Data parallelism: every replica starts from rank `src`'s fp32 master weights (plumbing over
        torch.distributed; the 16-bit operands are re-derived locally)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return
        with torch.cuda.device(self.device):
            dist.broadcast(self.params, src=src)
            self.repack()

    # ------------------------------------------------------------------ forward
    def workspace(self, B: int, T: int) -> _Workspace:
        key = (B, T)
        ws = self._workspaces.get(key)
        if ws is None or ws.stale():
            if len(self._workspaces) >= 64:  # (views only: the memory belongs to the arena)
                self._workspaces.pop(next(iter(self._workspaces)))
            generation = self.arena.generation
            with torch.cuda.device(self.device):
                ws = _Workspace(self, B, T)
            if self.arena.generation != generation:
                # a buffer was replaced by a larger one: cached workspaces that still look at the old storage
                # would keep it alive (a corpus whose batches keep getting longer must not pile up arenas)
                self._workspaces = {k: w for k, w in self._workspaces.items() if not w.stale()}
            self._workspaces[key] = ws
        return ws

    def _bind(self, ws: _Workspace) -> None:
        """`ws` takes over the arena buffers on the compute stream (called where a step starts: packing)."""
        if self._bound is not ws:
            ws.zero_allocation_rows()
            self._bound = ws

    def _pack_input(self, ws: _Workspace, x: torch.Tensor, training: bool) -> None:
        """fp32 (B,T,F) on the device -> packed bf16 operand of the first layer."""
        first = self.layers[0]
        self._bind(ws)
        if first.windowed:
            drop = training and self.dropout is not None and 0 in self.dropout_layers
            # forward(training=True) advances dropout_step before it derives the other layers' seeds
            seed = self._dropout_seed(0, self.dropout_step + 1) if drop else 0
            check(self.lib.sl_window_activation(ptr(x), ptr(ws.x_packed), ws.B, ws.T, first.cin, first.kernel,
                                                first.stride, first.cin_pad, self.precision,
                                                float(self.dropout) if drop else 0.0, seed, self.stream))
            ws.input_dropped = drop
        else:
            check(self.lib.sl_pack_activation(ptr(x), ptr(ws.x_packed), ws.B, ws.T, first.cin, ws.T_alloc,
                                              first.cin_pad, self.precision, self.stream))
        self.launches += 1

    def upload(self, input_batch, training: bool = False) -> _Workspace:
        """Host (B,T,F) float array (any float dtype, net.py:583) or a device fp32 tensor -> packed bf16.
        `training` only matters for raw-wave input, whose input dropout is applied while packing."""
        with torch.cuda.device(self.device):
            if isinstance(input_batch, torch.Tensor) and input_batch.is_cuda:
                B, T, F = input_batch.shape
                ws = self.workspace(B, T)
                x = input_batch.to(torch.float32).contiguous()
            elif isinstance(input_batch, torch.Tensor) and input_batch.is_pinned() \
                    and input_batch.dtype == torch.float32 and input_batch.is_contiguous():
                B, T, F = input_batch.shape  # already in pinned host memory: one async H2D copy
                ws = self.workspace(B, T)
                ws.x_f32.copy_(input_batch, non_blocking=True)
                x = ws.x_f32
            else:
                array = np.asarray(input_batch)
                B, T, F = array.shape
                if F != self.layers[0].cin:
                    raise ValueError("expected {} features per time step, got {}".format(self.layers[0].cin, F))
                ws = self.workspace(B, T)
                ws.x_host.copy_(torch.from_numpy(np.ascontiguousarray(array, dtype=np.float32)))
                ws.x_f32.copy_(ws.x_host, non_blocking=True)
                x = ws.x_f32
            first = self.layers[0]
            if F != first.cin:
                raise ValueError("expected {} features per time step, got {}".format(first.cin, F))
            self._pack_input(ws, x, training)
            self._current = ws
            return ws

    def stage_async(self, input_batch, slot: int) -> _Workspace:
        """Start the host->device copy of a (B,T,F) batch into input slot `slot` on the copy stream
        (pinned torch tensors go straight to the DMA engine; numpy arrays are first gathered into a
        pinned staging buffer by the calling host thread)."""
        with torch.cuda.device(self.device):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            if isinstance(input_batch, torch.Tensor):
                B, T, F = input_batch.shape
            else:
                input_batch = np.asarray(input_batch)
                B, T, F = input_batch.shape
            if F != self.layers[0].cin:
                raise ValueError("expected {} features per time step, got {}".format(self.layers[0].cin, F))
            ws = self.workspace(B, T)
            pinned = isinstance(input_batch, torch.Tensor) and input_batch.is_pinned() and \
                input_batch.dtype == torch.float32 and input_batch.is_contiguous()
            if not pinned:
                if self._slot_copied[slot] is not None:
                    self._slot_copied[slot].synchronize()  # the previous DMA out of this staging buffer is done
                source = input_batch if isinstance(input_batch, torch.Tensor) else torch.from_numpy(
                    np.ascontiguousarray(input_batch, dtype=np.float32))
                ws.host_slots[slot].copy_(source)
                input_batch = ws.host_slots[slot]
            copy = self._copy_stream
            if self._slot_consumed[slot] is not None:
                copy.wait_event(self._slot_consumed[slot])
            with torch.cuda.stream(copy):
                ws.x_slots[slot].copy_(input_batch, non_blocking=True)
                self._slot_copied[slot] = copy.record_event()
            return ws

    def consume_slot(self, ws: _Workspace, slot: int, training: bool = False) -> _Workspace:
        """Compute stream: wait for the slot's copy, pack it to bf16 and release the slot."""
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            main.wait_event(self._slot_copied[slot])
            self._pack_input(ws, ws.x_slots[slot], training)
            self._slot_consumed[slot] = main.record_event()
            self._current = ws
            return ws

    def _dropout_seed(self, index: int, step: Optional[int] = None) -> int:
        step = self.dropout_step if step is None else step
        mixed = (self.dropout_seed * 0x9E3779B97F4A7C15 + step * 0xD1B54A32D192ED03 +
                 (index + 1) * 0x8CB92BA72F3D8DD7) & 0xFFFFFFFFFFFFFFFF
        return mixed

    def forward(self, ws: Optional[_Workspace] = None, want_logits: bool = False,
                training: bool = False) -> _Workspace:
        """`training=True` is Keras' learning phase 1: dropout (if configured) is active."""
        ws = ws or self._current
        drop = training and self.dropout is not None
        if drop:
            self.dropout_step += 1
        with torch.cuda.device(self.device):
            x, t_in, t_alloc = ws.x_packed, ws.T0, ws.T_alloc
            for index, layer in enumerate(self.layers):
                if drop and layer.windowed and index in self.dropout_layers and not ws.input_dropped:
                    raise RuntimeError("raw-wave input dropout is applied while packing: upload(training=True)")
                if drop and index in self.dropout_layers and not layer.windowed:
                    if index not in ws.xdrop:
                        # (the dropout kernel writes the T_alloc padding rows as zeros itself)
                        ws.xdrop[index] = ws.take("xdrop{}".format(index), tuple(x.shape), x.dtype)
                        ws.bwd_mask[index] = ws.take("bwd_mask{}".format(index), (ws.B, t_in, layer.cin_pad // 8),
                                                     torch.uint8)
                    relu_below = ws.masks[index - 1] if index > 0 and self.layers[index - 1].activation == "relu" \
                        else None
                    seed = self._dropout_seed(index)
                    ws.dropout_seeds[index] = seed
                    check(self.lib.sl_dropout_fwd(ptr(x), ptr(ws.xdrop[index]), ptr(relu_below),
                                                  ptr(ws.bwd_mask[index]), ws.B, t_in, t_alloc, layer.gemm_cin,
                                                  self.precision,
                                                  float(self.dropout), seed, self.stream))
                    self.launches += 1
                    x = ws.xdrop[index]
                ws.layer_inputs[index] = x
                bias = self._b(self.params, layer)
                if layer.activation == "softmax":
                    self._timed("fwd", layer.name, lambda: self.lib.sl_conv1d_fwd(
                        ptr(x), ptr(self.w_fwd[index]), ptr(bias), None, None, ptr(ws.probs),
                        ptr(ws.logits) if want_logits else None, ptr(ws.logp), ws.B, t_in, t_alloc, 0,
                        layer.gemm_cin, layer.cout, layer.gemm_kernel, layer.gemm_stride, ACT_SOFTMAX, self.precision,
                        self.stream))
                else:
                    y = ws.acts[index]
                    act = ACT_RELU if layer.activation == "relu" else ACT_NONE
                    mask = ws.masks[index] if layer.activation == "relu" else None
                    self._timed("fwd", layer.name, lambda: self.lib.sl_conv1d_fwd(
                        ptr(x), ptr(self.w_fwd[index]), ptr(bias), ptr(y), ptr(mask), None, None, None, ws.B, t_in,
                        t_alloc, ws.t_alloc_out[index], layer.gemm_cin, layer.cout, layer.gemm_kernel,
                        layer.gemm_stride, act, self.precision, self.stream))
                    x, t_in, t_alloc = y, ws.t_out[index], ws.t_alloc_out[index]
                self.launches += 1
        return ws

    # ------------------------------------------------------------------ CTC
    def set_labels(self, ws: _Workspace, label_batch: np.ndarray, prediction_lengths: Sequence[int],
                   label_lengths: Sequence[int]) -> None:
        """label_batch (B, L_max) int32 padded with -1 (grapheme_enconding.py:25-32)."""
        label_batch = np.ascontiguousarray(label_batch, dtype=np.int32)
        if label_batch.ndim != 2 or label_batch.shape[0] != ws.B:
            raise ValueError("label batch must have shape (batch, max label length)")
        if label_batch.shape[1] == 0:
            label_batch = -np.ones((ws.B, 1), dtype=np.int32)
        V = self.layers[-1].cout
        pl = np.asarray(prediction_lengths, dtype=np.int64).reshape(-1)
        ll = np.asarray(label_lengths, dtype=np.int64).reshape(-1)
        for b in range(ws.B):
            label = label_batch[b, :ll[b]]
            if pl[b] > ws.Tp or pl[b] < 1:
                raise ValueError("prediction length {} outside [1, {}]".format(pl[b], ws.Tp))
            if label.size and (label.min() < 0 or label.max() >= V - 1):
                raise ValueError("label of example {} contains a grapheme outside [0, {})".format(b, V - 1))
            need = ctc_required_frames(label.tolist())
            if pl[b] < need:
                raise ValueError("Not enough time for target transition sequence "
                                 "(required: {}, available: {}) in example {}".format(need, pl[b], b))
        with torch.cuda.device(self.device):
            ws.labels = torch.from_numpy(label_batch).to(self.device)
            ws.input_len.copy_(torch.from_numpy(pl.astype(np.int32)))
            ws.label_len.copy_(torch.from_numpy(ll.astype(np.int32)))
            need_bytes = self.lib.sl_ctc_workspace_bytes(ws.B, ws.Tp, ws.labels.shape[1])
            if ws.ctc_ws is None or ws.ctc_ws.numel() < need_bytes:
                ws.ctc_ws = ws.take("ctc_ws", (need_bytes,), torch.uint8)

    def set_prediction_lengths(self, ws: _Workspace, prediction_lengths: Sequence[int]) -> None:
        pl = np.asarray(prediction_lengths, dtype=np.int32).reshape(-1)
        with torch.cuda.device(self.device):
            ws.input_len.copy_(torch.from_numpy(pl))

    def ctc(self, ws: _Workspace, want_grad: bool, grad_scale: float = 1.0, want_f32_grad: bool = False) -> torch.Tensor:
        """Per-utterance loss (device, (B,)); optionally d(grad_scale*sum loss)/d logits as the packed
        16-bit tile (times `ws.loss_scale` in the fp16 mode, see __init__)."""
        V = self.layers[-1].cout
        # |dlogits| <= grad_scale (softmax minus posterior): largest power of two S with grad_scale*S <= target
        ws.loss_scale = 2.0 ** math.floor(math.log2(self.loss_scale_target / grad_scale)) \
            if (want_grad and self.loss_scale_target > 0 and grad_scale > 0) else 1.0
        with torch.cuda.device(self.device):
            if want_grad:
                ws.ensure_backward(self)
                if want_f32_grad and ws.dz_f32 is None:
                    ws.dz_f32 = ws.take("dz_f32", (ws.B, ws.Tp, V), torch.float32)
            self._timed("ctc", "ctc_loss", lambda: self.lib.sl_ctc_loss(
                ptr(ws.logp), ptr(ws.probs), ptr(ws.labels), ptr(ws.input_len), ptr(ws.label_len), ptr(ws.loss),
                ptr(ws.dz_packed) if want_grad else None,
                ptr(ws.dz_f32) if (want_grad and want_f32_grad) else None, float(grad_scale * ws.loss_scale), ws.B,
                ws.Tp, V, ws.labels.shape[1], V - 1, self.precision, ptr(ws.ctc_ws), ws.ctc_ws.numel(), self.stream))
            self.launches += 2 if want_grad else 1
            if want_grad and want_f32_grad and ws.loss_scale != 1.0:
                ws.dz_f32.mul_(1.0 / ws.loss_scale)  # (parity tests only) report the unscaled gradient
        return ws.loss

    def greedy_decode(self, ws: _Workspace, merge_repeated: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        V = self.layers[-1].cout
        with torch.cuda.device(self.device):
            check(self.lib.sl_ctc_greedy_decode(ptr(ws.probs), ptr(ws.input_len), ptr(ws.decoded), ptr(ws.decoded_len),
                                                ws.B, ws.Tp, V, V - 1, 1 if merge_repeated else 0, self.stream))
            self.launches += 1
        return ws.decoded, ws.decoded_len

    def beam_search_decode(self, ws: _Workspace, beam_width: int = 100, top_paths: int = 1,
                           merge_repeated: bool = False, language_model=None
                           ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Prefix beam search over the softmax output of `forward` (tf.nn.ctc_beam_search_decoder, reference
        net.py:444-451): with the stock scorer, or — `language_model`: a `DeviceLanguageModel` — with the word
        n-gram scorer inside the search.  Returns (decoded (B, top_paths, T') int32 padded with -1,
        lengths (B, top_paths), log-probabilities (B, top_paths)), best path first."""
        V = self.layers[-1].cout
        with torch.cuda.device(self.device):
            need = self.lib.sl_ctc_beam_search_workspace_bytes(ws.B, ws.Tp, beam_width)
            if getattr(ws, "beam_ws", None) is None or ws.beam_ws.numel() < need:
                ws.beam_ws = ws.take("beam_ws", (need,), torch.uint8)
            decoded = torch.empty((ws.B, top_paths, ws.Tp), dtype=torch.int32, device=self.device)
            lengths = torch.empty((ws.B, top_paths), dtype=torch.int32, device=self.device)
            log_probabilities = torch.empty((ws.B, top_paths), dtype=torch.float32, device=self.device)
            if language_model is not None:
                import ctypes
                check(self.lib.sl_ctc_beam_search_decode_lm(
                    ptr(ws.probs), ptr(ws.input_len), ptr(decoded), ptr(lengths), ptr(log_probabilities), ws.B, ws.Tp,
                    V, V - 1, beam_width, top_paths, 1 if merge_repeated else 0, 1,
                    ctypes.addressof(language_model.struct), ptr(ws.beam_ws), ws.beam_ws.numel(), self.stream))
            else:
                check(self.lib.sl_ctc_beam_search_decode(ptr(ws.probs), ptr(ws.input_len), ptr(decoded), ptr(lengths),
                                                         ptr(log_probabilities), ws.B, ws.Tp, V, V - 1, beam_width,
                                                         top_paths, 1 if merge_repeated else 0, 1, ptr(ws.beam_ws),
                                                         ws.beam_ws.numel(), self.stream))
            self.launches += 1
        return decoded, lengths, log_probabilities

    # ------------------------------------------------------------------ backward + update
    def ensure_training_state(self) -> None:
        with torch.cuda.device(self.device):
            if self.grads is None:
                self.grads = torch.zeros_like(self.params)
                self.adam_m = torch.zeros_like(self.params)
                self.adam_v = torch.zeros_like(self.params)

    def grad_buckets(self) -> List[Tuple[int, int, int]]:
        """(first layer index, begin, end) float ranges of the flat gradient buffer, in the order
        backward completes them: [output_conv + big_conv_2], [big_conv_1], [the inner layers], [the first
        trainable layer] (16 / 64 / 12 / 6 MB at reference widths).  The last bucket is the only
        one whose all-reduce and Adam update nothing is left to hide behind, so it is kept to the one layer whose
        weight gradient finishes last (SL_BUCKET_SPLIT_FIRST=0: one 18 MB bucket for everything below big_conv_1)."""
        n = len(self.layers)
        first = self.first_trainable()
        cuts = {max(first, n - 2), max(first, n - 3), first}
        if os.environ.get("SL_BUCKET_SPLIT_FIRST", "1") != "0" and first + 1 < n - 3:
            cuts.add(first + 1)
        cuts = sorted(cuts)
        cuts = [c for c in cuts if c < n]
        buckets = []
        end_layer = n
        for start_layer in reversed(cuts):
            if start_layer >= end_layer:
                continue
            begin = self.layers[start_layer].w_offset
            last = self.layers[end_layer - 1]
            buckets.append((start_layer, begin, last.b_offset + last.cout_pad))
            end_layer = start_layer
        return buckets

    def backward(self, ws: Optional[_Workspace] = None, on_bucket_ready=None, on_bucket_consumed=None) -> None:
        """Fill self.grads from ws.dz_packed (set by ctc(want_grad=True)).

        `on_bucket_ready(begin, end)` is called (on the launching stream) as soon as the weight
        gradients of a bucket of `grad_buckets()` have been enqueued, so the data-parallel
        all-reduce of the big top layers overlaps the backward pass of the layers below.
        `on_bucket_consumed(begin, end)` follows once the input gradient of the bucket's lowest
        layer — the last reader of the bucket's 16-bit weights — has been enqueued too: from that
        point of the stream on, the bucket's parameters may be updated.

        The chain dY_l -> dgrad_l -> dY_{l-1} runs on the current stream; with
        `overlap_backward` the weight gradients (which only consume dY_l and the saved
        activations) run on a side stream, so the partial last wave of one persistent kernel
        is filled by CTAs of the other."""
        ws = ws or self._current
        self.ensure_training_state()
        first = self.first_trainable()
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            side = None
            if self.overlap_backward and self.profile is None:
                if self._side_stream is None:
                    self._side_stream = torch.cuda.Stream(device=self.device)
                side = self._side_stream
            if not self._grads_clean:
                self.grads.zero_()
            self._grads_clean = False
            dy = ws.dz_packed
            flip = 0
            wgrad_done = None  # event: the wgrad that still reads the buffer the next dgrad overwrites
            bucket_starts = {b[0]: (b[1], b[2]) for b in self.grad_buckets()} \
                if (on_bucket_ready or on_bucket_consumed) else {}
            for index in range(len(self.layers) - 1, first - 1, -1):
                layer = self.layers[index]
                x = ws.layer_inputs[index]  # the tensor the forward conv actually read (dropped or not)
                t_in = ws.T0 if index == 0 else ws.t_out[index - 1]
                t_alloc = ws.T_alloc if index == 0 else ws.t_alloc_out[index - 1]
                launch_wgrad = lambda: self._timed("wgrad", layer.name, lambda: self.lib.sl_conv1d_wgrad(
                    ptr(x), ptr(dy), ptr(self._w(self.grads, layer)), ptr(self._b(self.grads, layer)), ws.B, t_in,
                    t_alloc, layer.gemm_cin, layer.cout, layer.gemm_kernel, layer.gemm_stride, self.precision, 1,
                    1.0 / ws.loss_scale, self.stream))
                previous_wgrad_done = wgrad_done
                self._lift_sm_limit()
                if side is None:
                    launch_wgrad()
                else:
                    side.wait_event(main.record_event())
                    with torch.cuda.stream(side):
                        launch_wgrad()
                        wgrad_done = side.record_event()
                self.launches += 1
                if index in bucket_starts and on_bucket_ready is not None:
                    if side is not None:
                        main.wait_stream(side)
                    on_bucket_ready(*bucket_starts[index])
                if index in bucket_starts and on_bucket_consumed is not None and index == first:
                    on_bucket_consumed(*bucket_starts[index])  # the lowest trainable layer has no input gradient
                if index > first:
                    below = self.layers[index - 1]
                    dx = ws.dact[flip].view(-1)[:ws.B * t_in * self.planes * below.cout_pad].view(
                        ws.B, t_in, self.planes * below.cout_pad)
                    dropped = x is ws.xdrop.get(index)
                    if dropped:  # keep & ReLU bits written by the dropout kernel; scale 1/(1-p)
                        mask, out_scale = ws.bwd_mask[index], self.dropout_scale
                    else:
                        mask, out_scale = (ws.masks[index - 1] if below.activation == "relu" else None), 1.0
                    if previous_wgrad_done is not None:
                        main.wait_event(previous_wgrad_done)  # it reads the buffer dx aliases
                    need = 0 if layer.gemm_stride != 1 else self.lib.sl_conv1d_dgrad_workspace_bytes(
                        ws.B, t_in, layer.gemm_cin, layer.cout, layer.gemm_kernel)
                    if need and (ws.dgrad_ws is None or ws.dgrad_ws.numel() < need):
                        ws.dgrad_ws = ws.take("dgrad_ws", (need,), torch.uint8)
                    scratch = ws.dgrad_ws if need else None
                    self._lift_sm_limit()
                    self._timed("dgrad", layer.name, lambda: self.lib.sl_conv1d_dgrad(
                        ptr(dy), ptr(self.w_fwd[index]), ptr(mask), ptr(dx), ws.B, t_in, layer.gemm_cin, layer.cout,
                        layer.gemm_kernel, layer.gemm_stride, self.precision, out_scale, ptr(scratch), need,
                        self.stream))
                    self.launches += 2 if (need or layer.gemm_stride == 2) else 1
                    if index in bucket_starts and on_bucket_consumed is not None:
                        on_bucket_consumed(*bucket_starts[index])
                    dy = dx
                    flip ^= 1
            if side is not None:
                main.wait_stream(side)

    def adam_step(self, lr: float, beta_1: float, beta_2: float, epsilon: float, iteration: int) -> None:
        """Keras-2 Adam over the whole flat buffer (frozen layers have zero gradients and zero
        moments, so they do not move) fused with the refresh of the bf16 operands."""
        import ctypes
        if self._adam_tables is None:
            n = len(self.layers)
            begins = (ctypes.c_size_t * n)(*[l.w_offset for l in self.layers])
            ends = (ctypes.c_size_t * n)(*[l.w_offset + l.w_size for l in self.layers])
            targets = (ctypes.c_void_p * n)(*[w.data_ptr() for w in self.w_fwd])
            cin_pads = (ctypes.c_int * n)(*[l.cin_pad for l in self.layers])
            self._adam_tables = (begins, ends, targets, cin_pads)
        begins, ends, targets, cin_pads = self._adam_tables
        with torch.cuda.device(self.device):
            self._timed("adam", "adam", lambda: self.lib.sl_adam_step_fused(
                ptr(self.params), ptr(self.grads), ptr(self.adam_m), ptr(self.adam_v), self.param_count, begins, ends,
                targets, cin_pads, len(self.layers), self.precision, lr, beta_1, beta_2, epsilon, iteration,
                self.stream))
            self.launches += 1

    def adam_step_range(self, begin: int, end: int, lr: float, beta_1: float, beta_2: float, epsilon: float,
                        iteration: int, stream: Optional[int] = None) -> None:
        """The same fused update on floats [begin, end) of the flat buffer only (a bucket of
        `grad_buckets()`: whole layers), on `stream`."""
        import ctypes
        inside = [(i, l) for i, l in enumerate(self.layers) if begin <= l.w_offset and l.w_offset + l.w_size <= end]
        n = len(inside)
        begins = (ctypes.c_size_t * max(n, 1))(*[l.w_offset - begin for _, l in inside])
        ends = (ctypes.c_size_t * max(n, 1))(*[l.w_offset + l.w_size - begin for _, l in inside])
        targets = (ctypes.c_void_p * max(n, 1))(*[self.w_fwd[i].data_ptr() for i, _ in inside])
        cin_pads = (ctypes.c_int * max(n, 1))(*[l.cin_pad for _, l in inside])
        offset = 4 * begin
        with torch.cuda.device(self.device):
            self._timed("adam", "adam", lambda: self.lib.sl_adam_step_fused(
                self.params.data_ptr() + offset, self.grads.data_ptr() + offset, self.adam_m.data_ptr() + offset,
                self.adam_v.data_ptr() + offset, end - begin, begins, ends, targets, cin_pads, n, self.precision, lr,
                beta_1, beta_2, epsilon, iteration, self.stream if stream is None else stream))
            self.launches += 1

    def backward_and_update(self, ws: Optional[_Workspace], loss: torch.Tensor, optimizer, data_parallel=None
                            ) -> torch.Tensor:
        """Backward pass, the data-parallel exchange step and the Keras-2 Adam update, pipelined by
        gradient bucket: as soon as the weight gradients of a bucket exist, an UPDATE
        STREAM all-reduces them over NVLink (few CTAs: `DataParallel.max_ctas`) and applies Adam to that
        bucket, while the compute stream carries on with the input / weight gradients of the layers
        below.  Only the last, small bucket's all-reduce + update are exposed.  Without data parallelism
        the same pipeline hides the optimizer step behind the backward pass.

        While a bucket's all-reduce is in flight the persistent Conv1D grids of the next
        `data_parallel.limited_launches` launches are sized to 148 - max_ctas CTAs
        (`sl_set_sm_limit`), so the collective finds free SMs instead of delaying the tail of a
        148-CTA grid.  `optimizer.iterations` must already count this step.  Returns the (global)
        sum of the per-utterance losses as a device scalar."""
        ws = ws or self._current
        dp = data_parallel if (data_parallel is not None and data_parallel.active) else None
        hyper = (optimizer.lr, optimizer.beta_1, optimizer.beta_2, optimizer.epsilon, optimizer.iterations)
        self.ensure_training_state()
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            loss_sum = loss.sum()
            if self.profile is not None or not self.pipeline_update:
                # measurement / A-B mode: everything in order on the compute stream
                self.backward(ws, on_bucket_ready=(lambda b, e: dp.allreduce_range(self.grads, b, e)) if dp else None)
                if dp:
                    dp.allreduce_scalar(loss_sum)
                for _, begin, end in self.grad_buckets():
                    self.adam_step_range(begin, end, *hyper)
                self._clear_gradients(main)
                return loss_sum
            if self._update_stream is None:
                self._update_stream = torch.cuda.Stream(device=self.device)
            update = self._update_stream
            update.wait_stream(main)  # (the previous step's update has long been joined; orders the zero fill)

            def ready(begin, end):
                if dp is None:
                    return
                update.wait_event(main.record_event())
                dp.allreduce_range(self.grads, begin, end, stream=update)
                if dp.max_ctas > 0 and dp.limited_launches > 0:
                    check(self.lib.sl_set_sm_limit(max(1, self.sm_count - dp.max_ctas)))
                    self._sm_limit_until = self.launches + dp.limited_launches

            def consumed(begin, end):
                update.wait_event(main.record_event())
                self.adam_step_range(begin, end, *hyper, stream=update.cuda_stream)
                self._clear_gradients(update, begin, end)

            try:
                self.backward(ws, on_bucket_ready=ready, on_bucket_consumed=consumed)
            finally:
                self._lift_sm_limit(force=True)
            if dp:
                dp.allreduce_scalar(loss_sum, stream=update, after=main)
            main.wait_stream(update)
            self._grads_clean = not self.retain_gradients and self.first_trainable() == 0
        return loss_sum

    def _clear_gradients(self, stream, begin: int = 0, end: Optional[int] = None) -> None:
        """Zero grads[begin:end] on `stream` once the optimizer has consumed them (see `retain_gradients`)."""
        if self.retain_gradients:
            return
        with torch.cuda.stream(stream):
            self.grads[begin:end].zero_()
        if end is None:
            self._grads_clean = self.first_trainable() == 0

    def _lift_sm_limit(self, force: bool = False) -> None:
        """Back to full-width persistent grids once the launches that overlap an all-reduce are out."""
        if self._sm_limit_until and (force or self.launches >= self._sm_limit_until):
            check(self.lib.sl_set_sm_limit(0))
            self._sm_limit_until = 0

    def sync(self) -> None:
        torch.cuda.current_stream(self.device).synchronize()
