#!/usr/bin/env python
"""bench.py — spectrogram frames/sec of the wav2letter hot path (fwd + CTC + bwd + all-reduce + Adam).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload full|small|long|ragged] [--batch-per-gpu B] [--dtype fp16|bf16|bf16x2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  Metric / workloads: BASELINE.json.  A "step" is one
training step on one synthetic batch of B utterances per GPU (weak scaling):
  value  — frames/s with the batch already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e    — frames/s through the public `Wav2Letter.fit_batches` call with HOST inputs
           (pinned host -> device copy of the batch and device -> host read of the loss inside
           the timed region)
  roofline — the dominant kernel (largest share of the step) against MEASURED_PEAKS.json
  precision — the mode's measured error against the fp64 oracle (logits, gradients) on a small
           reference-width case run inside this process, next to BASELINE.json's tolerance
  twins  — the same workload and batch in the other precision modes (ms/step, frames/s, conv roofline)
  configs — short runs of the other BASELINE.json configurations (small / long-form / ragged)
  dp_check — (N > 1) gradients of the data-parallel step against the single-GPU step on the
           global batch, replicas identical
  cpu_baseline — the torch-CPU restatement of the Keras/TF path (Keras/TF are not installable
           offline) on a bounded sample, all host cores (N=1, rank 0 only)
`--impl reference` times that CPU restatement alone, on the same config / metric / unit.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (main filters, out filters, seconds, default batch per GPU, default dtype, BASELINE.json config)
    "full": (250, 2000, 10.0, 64, "fp16", "wav2letter-full (250/2000 filters) batch=64/GPU, 10 s utterances"),
    "small": (256, 256, 10.0, 32, "bf16x2", "wav2letter-small (11 Conv1D, 256ch, mel-128) batch=32, 10 s, fp32-parity"),
    "long": (250, 2000, 60.0, 16, "fp16", "long-form 60 s utterances (T=7501, T'=3751), CTC stress"),
    # durations U(2 s, 17 s), zero padded to the batch maximum, padding NOT masked
    "ragged": (250, 2000, None, 64, "fp16", "LibriSpeech-shaped ragged batch: U(2 s, 17 s) utterances padded to the "
                                            "batch maximum (unmasked padding, net.py:578-587), batch=64/GPU"),
}
METRIC = "spectrogram frames/sec, fwd+bwd+CTC+Adam training step"
DATA = "synthetic N(0,1) mel-128 spectrograms, uniform random labels, glorot-uniform weights"
TOLERANCE = {"logits_rel": 1e-3, "ctc_loss_rel": 1e-4}  # BASELINE.json north_star


def conv_flops(layers, t_outs, batch, first_trainable=0):
    """Algorithmic FLOPs (un-padded channels, MAC = 2) per kernel kind and layer."""
    out = {}
    for index, (layer, t_out) in enumerate(zip(layers, t_outs)):
        f = 2.0 * layer.kernel * layer.cin * layer.cout * t_out * batch
        out[("fwd", layer.name)] = f
        if index >= first_trainable:
            out[("wgrad", layer.name)] = f
        if index > first_trainable:
            out[("dgrad", layer.name)] = f
    return out


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Start of the timed region: nvidia-smi takes ~0.2 s to produce its first line, so the sampler is
        started before the warm-up and only the lines read from here on are counted."""
        deadline = time.perf_counter() + 3.0
        while self.proc is not None and not self.lines and time.perf_counter() < deadline:
            time.sleep(0.02)  # (before the timed region starts)
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, sm_max, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        for stamp, line in self.lines:
            if stamp < t_mark:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max = float(parts[2])
            except ValueError:
                continue
            try:
                power.append(float(parts[3]))
            except ValueError:
                pass
            for name, value in zip(names, parts[4:8]):
                if value.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        power.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": sm_max, "samples": len(sm),
                "reasons": sorted(reasons), "power_w": power[len(power) // 2] if power else None}


def measured_peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        data = json.loads(path.read_text())
        return {"hbm_gbs": data["hbm_gbs"], "bf16_tflops": data["bf16_tflops"],
                "bf16_tflops_sustained": data.get("bf16_tflops_sustained", data["bf16_tflops"]), "source": "measured"}
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(kernel_key, workload, dtype):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the named kernel from
    the committed `ncu --set full` capture summary (profiles/ncu_traffic.json), or None.  fp16 and bf16
    move the same bytes (both one 16-bit plane), so fp16 falls back to the bf16 capture."""
    path = ROOT / "profiles" / "ncu_traffic.json"
    if not path.exists():
        return None
    table = json.loads(path.read_text())
    for key_dtype in (dtype, "bf16" if dtype == "fp16" else dtype):
        entry = table.get("{}|{}|{}".format(workload, key_dtype, kernel_key))
        if entry:
            return entry["dram_bytes"]
    return None


def workload_frames(workload, batch, seed):
    """Per-utterance frame counts of one synthetic batch."""
    import numpy as np
    from speechless_b200.synthetic import frames_for_seconds
    seconds = WORKLOADS[workload][2]
    if seconds is not None:
        return [frames_for_seconds(seconds)] * batch
    rng = np.random.default_rng(977 + seed)
    return [frames_for_seconds(float(s)) for s in rng.uniform(2.0, 17.0, size=batch)]


def make_examples(workload, batch, rank):
    from speechless_b200 import english_frequent_characters as alphabet
    from speechless_b200.synthetic import synthetic_batch
    frames = workload_frames(workload, batch, rank)
    return synthetic_batch(batch, frames, alphabet, seed=1234 + rank), frames, alphabet


def config_record(workload, batch, world, frames, label_length):
    return {"workload": WORKLOADS[workload][5], "global_batch": batch * world, "batch_per_gpu": batch,
            "frames_per_utterance": max(frames), "valid_frames_per_gpu_batch": int(sum(frames)),
            "label_length": label_length, "parallelism": "dp{}".format(world),
            "l2": "per-step working set (~1 GB of activations) exceeds the 126 MB L2; no explicit flush"}


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the torch-CPU restatement of the Keras path (oracle/torch_cpu.py)
# --------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(workload, steps, warmup, sample_batch, budget_seconds):
    """frames/s of the torch-CPU restatement on a bounded sample of the workload: `sample_batch` utterances
    of the workload's own lengths per step, exactly `warmup` + `steps` steps unless the budget runs out."""
    import numpy as np
    import torch
    from oracle.torch_cpu import TorchCpuWav2Letter
    from speechless_b200.grapheme_enconding import CtcGraphemeEncoding
    main, out = WORKLOADS[workload][0], WORKLOADS[workload][1]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch, frames, alphabet = make_examples(workload, sample_batch, 0)
    encoding = CtcGraphemeEncoding(alphabet)
    T = max(frames)
    x = np.zeros((sample_batch, T, 128), dtype=np.float32)
    for i, e in enumerate(batch):
        x[i, :frames[i]] = e.z_normalized_transposed_spectrogram()
    labels = encoding.encode_label_batch([e.label for e in batch])
    pred_len = [f // 2 for f in frames]
    label_len = [len(e.label) for e in batch]
    model = TorchCpuWav2Letter(128, len(alphabet) + 1, main, out, seed=0)
    started = time.perf_counter()
    for _ in range(warmup):
        model.train_step(x, labels, pred_len, label_len)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        model.train_step(x, labels, pred_len, label_len)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - started > budget_seconds:
            break
    mean = sum(times) / len(times)
    # the metric counts valid frames, sum_b T_b (the step also computes on the padding, as the Keras path does)
    return {"value": sum(frames) / mean, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": "torch-CPU fp32 restatement of the Keras/TF path (Keras/TF not installable offline); "
                      "{} utterances x {} frames per step, {} warm-up + {} timed step(s), {:.2f} s/step".format(
                sample_batch, T, warmup, len(times), mean)}, mean, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    main, out, seconds, default_batch, default_dtype, config_name = WORKLOADS[args.workload]
    batch = args.batch_per_gpu or default_batch
    examples, frames, alphabet = make_examples(args.workload, batch, 0)
    # bounded sample: 16 utterances of the workload per step (the CPU path's rate per frame does not depend on
    # the batch beyond that: oneDNN already uses every core), fewer only when K + W steps would not fit ~4 min
    per_utt_seconds = max(frames) / 1251.0 * 0.1  # ~0.1 s per 10 s utterance on 16 cores (profiles/, round 1)
    sample = 16
    while sample > 2 and (args.steps + args.warmup) * sample * per_utt_seconds > 240.0:
        sample //= 2
    sample = min(sample, batch)
    baseline, mean, timed = cpu_reference_rate(args.workload, args.steps, args.warmup, sample, budget_seconds=420.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": baseline["value"], "unit": "frames/s", "n_gpus": args.gpus,
        "steps": timed, "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": DATA,
        "config": config_record(args.workload, batch, world, frames, len(examples[0].label)),
        "cpu_baseline": baseline,
        "e2e": {"value": baseline["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
class Arm:
    """One (workload, batch, dtype) instance on this rank: model, HBM-resident batch, step function."""

    def __init__(self, workload, batch, dtype, dp, device, overlap_backward=False):
        import torch
        from speechless_b200.net import Wav2Letter
        self.torch = torch
        self.workload, self.batch, self.dtype, self.dp = workload, batch, dtype, dp
        main, out = WORKLOADS[workload][0], WORKLOADS[workload][1]
        self.examples, self.frames, self.alphabet = make_examples(workload, batch, dp.rank)
        self.global_batch = batch * dp.world_size
        self.net = Wav2Letter(128, self.alphabet, main_filter_count=main, out_filter_count=out, compute_dtype=dtype,
                              device=device, seed=0)
        self.tower = self.net.tower
        self.tower.overlap_backward = bool(overlap_backward)
        self.inputs, _ = self.net._inputs_for_loss_net(self.examples)
        names = Wav2Letter.InputNames
        self.names = names
        self.host_x = torch.from_numpy(self.inputs[names.input_batch]).pin_memory()
        self.ws = self.tower.upload(self.host_x)
        self.tower.set_labels(self.ws, self.inputs[names.label_batch], self.inputs[names.prediction_lengths],
                              self.inputs[names.label_lengths])
        # the metric counts VALID frames, sum_b T_b over the global batch; the step also computes
        # on the zero padding of a ragged batch (the reference does not mask it): padded_frames_per_step
        self.frames_per_step = dp.world_size * sum(self.frames)
        self.padded_frames_per_step = self.global_batch * max(self.frames)

    def device_step(self):
        tower, net = self.tower, self.net
        ws = tower.upload(self.ws.x_f32)  # re-pack from the HBM-resident fp32 batch
        if ws is not self.ws:  # (the tower's arena grew under another shape: fresh views of the same buffers)
            tower.set_labels(ws, self.inputs[self.names.label_batch], self.inputs[self.names.prediction_lengths],
                             self.inputs[self.names.label_lengths])
            self.ws = ws
        tower.forward(ws)
        loss = tower.ctc(ws, want_grad=True, grad_scale=1.0 / self.global_batch)
        net.optimizer.iterations += 1
        tower.backward_and_update(ws, loss, net.optimizer, data_parallel=self.dp)
        return loss

    def ctc_lattices_only_ms(self, repetitions=20):
        """The alpha / beta lattice kernel alone (`sl_ctc_loss` without the gradient output) on the probabilities of
        the last step, CUDA-event timed back to back: what is left of the `ctc` time is the gradient kernel."""
        torch = self.torch
        for _ in range(3):
            self.tower.ctc(self.ws, want_grad=False)
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(repetitions):
            self.tower.ctc(self.ws, want_grad=False)
        stop.record()
        torch.cuda.synchronize()
        return start.elapsed_time(stop) / repetitions

    def timed(self, steps, warmup):
        """ms/step, device-timed with CUDA events, max over ranks."""
        torch, dp = self.torch, self.dp
        for _ in range(warmup):
            self.device_step()
        torch.cuda.synchronize()
        dp.barrier()
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            loss = self.device_step()
        stop.record()
        torch.cuda.synchronize()
        dp.barrier()
        torch.cuda.synchronize()
        return dp.max_over_ranks(start.elapsed_time(stop)) / steps, loss

    def instrumented(self, steps):
        """The same steps with a CUDA-event pair around every launch -> per-kernel ms per step."""
        torch = self.torch
        self.tower.profile = []
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            self.device_step()
        stop.record()
        torch.cuda.synchronize()
        self.dp.barrier()
        profile, self.tower.profile = self.tower.profile, None
        agg = {}
        for kind, name, ev0, ev1 in profile:
            agg[(kind, name)] = agg.get((kind, name), 0.0) + ev0.elapsed_time(ev1)
        # per step: a kernel kind launched several times per step (per-bucket Adam) adds up
        return {key: total / steps for key, total in agg.items()}, start.elapsed_time(stop) / steps

    def conv_summary(self, per_kernel_ms, peaks):
        flops = conv_flops(self.tower.layers, self.ws.t_out, self.batch, self.tower.first_trainable())
        conv_ms = sum(ms for key, ms in per_kernel_ms.items() if key in flops)
        total = sum(flops.values())
        achieved = total / (conv_ms * 1e-3) / 1e12
        return flops, {"achieved": round(achieved, 1), "frac": round(achieved / peaks["bf16_tflops"], 4),
                       "frac_sustained": round(achieved / peaks["bf16_tflops_sustained"], 4), "ms": round(conv_ms, 3),
                       "flops_per_step": total}


def per_kernel_ctc(per_kernel_ms):
    return sum(ms for (kind, _), ms in per_kernel_ms.items() if kind == "ctc")


def ctc_algorithmic_bytes(frames, examples, V=29):
    """SURVEY.md §8d, per utterance: read y + write grad (2 P V 4) + write/read alpha (2 P S 4) + labels."""
    return sum((f // 2) * (2 * V * 4 + 2 * (2 * len(e.label) + 1) * 4) + len(e.label) * 4
               for f, e in zip(frames, examples))


def varying_shapes_run(arm, steps, distinct=4):
    """Real corpora give every batch its own longest utterance: the public pipelined training call
    (`Wav2Letter.fit_batches`, host inputs) over `distinct` ragged batches of different padded lengths in turn —
    each step re-binds the per-shape workspace views of the tower's arena (no allocation after the first round)."""
    import torch
    from speechless_b200.synthetic import synthetic_batch
    net, dp = arm.net, arm.dp
    batches = []
    for k in range(distinct):
        frames = workload_frames(arm.workload, arm.batch, 1000 * (k + 1) + dp.rank)
        examples = synthetic_batch(arm.batch, frames, arm.alphabet, seed=77 + k + 10 * dp.rank)
        inputs, _ = net._inputs_for_loss_net(examples)
        inputs[arm.names.input_batch] = torch.from_numpy(inputs[arm.names.input_batch]).pin_memory()
        batches.append((inputs, sum(frames), max(frames)))

    def run(n):
        return net.fit_batches((batches[i % distinct][0] for i in range(n)), global_batch_size=arm.global_batch,
                               data_parallel=dp if dp.active else None)
    run(2 * distinct)  # first round allocates
    torch.cuda.synchronize()
    dp.barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    start.record()
    run(steps)
    stop.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms = dp.max_over_ranks(max(start.elapsed_time(stop), wall)) / steps
    valid = sum(batches[i % distinct][1] for i in range(steps)) / steps * dp.world_size
    return {"ms_per_step": round(ms, 4), "value": valid / (ms * 1e-3), "unit": "frames/s (valid frames, end to end)",
            "distinct_shapes": distinct, "padded_lengths": [b[2] for b in batches], "steps": steps}


def precision_record(dtype, device):
    """Measured error of `dtype` against the fp64 oracle on a small reference-width case (250/2000 filters,
    2 ragged utterances): logits, CTC loss, and the gradient of every kernel/bias — raw, and given the ReLU
    sign pattern the device saw (the gradient is discontinuous where a pre-activation crosses zero)."""
    import numpy as np
    from oracle import keras_tf_oracle as oracle
    from speechless_b200 import english_frequent_characters as alphabet
    from speechless_b200.net import Wav2Letter
    from speechless_b200.synthetic import synthetic_batch

    def rel(got, want):
        return float(np.abs(np.asarray(got, dtype=np.float64) - want).max() / max(np.abs(want).max(), 1e-30))

    net = Wav2Letter(128, alphabet, compute_dtype=dtype, device=device, seed=1)
    rng = np.random.default_rng(101)
    for layer in net.predictive_net.layers:
        kernel, bias = layer.get_weights()
        layer.set_weights([kernel, (rng.standard_normal(bias.shape) * 0.1).astype(np.float32)])
    ref = oracle.Wav2LetterOracle(128, len(alphabet) + 1, 250, 2000, dtype=np.float64)
    weights = [layer.get_weights() for layer in net.predictive_net.layers]
    ref.set_weights([w for w, _ in weights], [b for _, b in weights])
    batch = synthetic_batch(2, [203, 171], alphabet, seed=6, label_length=14)
    inputs, _ = net._inputs_for_loss_net(batch)
    names = Wav2Letter.InputNames
    tower = net.tower
    ws = tower.upload(inputs[names.input_batch])
    tower.forward(ws, want_logits=True)
    tower.set_labels(ws, inputs[names.label_batch], inputs[names.prediction_lengths], inputs[names.label_lengths])
    loss = tower.ctc(ws, want_grad=True, grad_scale=0.5)
    tower.backward(ws)
    tower.sync()
    logits, loss = ws.logits.cpu().numpy(), loss.cpu().numpy()
    saved = tower.params.clone()
    tower.params.copy_(tower.grads)
    grads = [tower.get_layer_weights(i) for i in range(len(tower.layers))]
    tower.params.copy_(saved)
    masks = {i: np.unpackbits(ws.masks[i].cpu().numpy(), axis=2, bitorder="little")[..., :l.cout].astype(bool)
             for i, l in enumerate(tower.layers[:-1])}
    x = inputs[names.input_batch].astype(np.float64)
    args = (inputs[names.label_batch], inputs[names.prediction_lengths][:, 0], inputs[names.label_lengths][:, 0])
    losses_ref, _, logits_ref, dws, dbs = ref.loss_and_gradients(x, *args)
    _, _, _, dws_m, dbs_m = ref.loss_and_gradients(x, *args, relu_masks=masks)
    record = {
        "mode": dtype, "checked_on": "250/2000 filters, 2 utterances (203 / 171 frames), fp64 numpy oracle, in this run",
        "logits_rel_err": rel(logits, logits_ref), "ctc_loss_rel_err": float(np.abs(loss / losses_ref - 1).max()),
        "grad_rel_err": max(max(rel(g[0], dws[i]), rel(g[1], dbs[i])) for i, g in enumerate(grads)),
        "grad_rel_err_given_relu_pattern": max(max(rel(g[0], dws_m[i]), rel(g[1], dbs_m[i]))
                                               for i, g in enumerate(grads)),
        "tolerance": TOLERANCE,
    }
    record["meets_tolerance"] = bool(record["logits_rel_err"] <= TOLERANCE["logits_rel"] and
                                     record["ctc_loss_rel_err"] <= TOLERANCE["ctc_loss_rel"])
    return record


def dp_equivalence(dp, device):
    """N > 1: the pipelined data-parallel step (per-bucket all-reduce + Adam) against the single-GPU step on
    the global batch, small widths; every rank computes both and the verdicts are gathered."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from speechless_b200 import english_frequent_characters as alphabet
    from speechless_b200.net import Wav2Letter
    from speechless_b200.synthetic import synthetic_batch
    global_batch = 4 * dp.world_size
    batches = [synthetic_batch(global_batch, [300 - 5 * i - 3 * s for i in range(global_batch)], alphabet, seed=5 + s,
                               label_length=20) for s in range(2)]
    kwargs = dict(main_filter_count=128, out_filter_count=256, seed=3, device=device, compute_dtype="bf16x2")
    single = Wav2Letter(128, alphabet, **kwargs)
    single.tower.retain_gradients = True  # (the step otherwise clears each gradient bucket after its update)
    single_losses = [single.train_on_batch(single._inputs_for_loss_net(b)[0]) for b in batches]
    net = Wav2Letter(128, alphabet, **kwargs)
    net.tower.retain_gradients = True
    losses = []
    for b in batches:
        longest = max(e.z_normalized_transposed_spectrogram().shape[0] for e in b)
        inputs = net._input_dictionary_for_loss_net(dp.shard(b), pad_to_length=longest)
        losses.append(net.train_on_batch(inputs, global_batch_size=global_batch, data_parallel=dp))
    torch.cuda.synchronize()
    reference = net.tower.params.clone()
    dist.broadcast(reference, src=0)
    scale = float(single.tower.grads.abs().max())
    mine = {"replicas_identical": bool(torch.equal(reference, net.tower.params)),
            "grad_rel_err": float((net.tower.grads - single.tower.grads).abs().max()) / scale,
            "param_abs_err": float((net.tower.params - single.tower.params).abs().max()),
            "loss_rel_err": float(np.abs(np.array(losses) / np.array(single_losses) - 1).max())}
    everyone = dp.gather_objects(mine)
    return {"replicas_identical": all(e["replicas_identical"] for e in everyone),
            "grad_rel_err": max(e["grad_rel_err"] for e in everyone),
            "param_abs_err": max(e["param_abs_err"] for e in everyone),
            "loss_rel_err": max(e["loss_rel_err"] for e in everyone),
            "checked_on": "2 pipelined data-parallel Adam steps vs the single-GPU steps on the global batch of {} "
                          "ragged utterances, 128/256 filters, bf16x2".format(global_batch),
            "own_communicator": dp.owns_communicator, "comm_max_ctas": dp.max_ctas,
            "limited_launches": dp.limited_launches}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from speechless_b200.distributed import DataParallel

    # stdout carries exactly one JSON line: when the caller asks NCCL for INFO / VERSION output it goes to
    # stderr (the caller's NCCL_DEBUG itself is left alone)
    if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    dp = DataParallel()
    rank, world = dp.rank, dp.world_size
    if world != args.gpus and rank == 0:
        print("warning: --gpus {} but WORLD_SIZE {}".format(args.gpus, world), file=sys.stderr)
    torch.cuda.set_device(dp.local_rank)
    device = torch.device("cuda", dp.local_rank)

    default_batch, default_dtype = WORKLOADS[args.workload][3], WORKLOADS[args.workload][4]
    dtype = args.dtype or default_dtype
    batch = args.batch_per_gpu or default_batch
    peaks = measured_peaks()

    arm = Arm(args.workload, batch, dtype, dp, device, overlap_backward=args.overlap_backward)
    tower, net = arm.tower, arm.net
    tower.pipeline_update = bool(args.pipeline_update)

    # ---- device-resident arm ----
    sampler = ClockSampler(dp.local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        arm.device_step()
    torch.cuda.synchronize()
    dp.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.mark()
    launches_before = tower.launches
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_t0 = time.perf_counter()
    start.record()
    for _ in range(args.steps):
        loss = arm.device_step()
    stop.record()
    host_enqueue_ms = (time.perf_counter() - host_t0) * 1e3 / args.steps
    torch.cuda.synchronize()
    dp.barrier()
    torch.cuda.synchronize()
    ms_device = dp.max_over_ranks(start.elapsed_time(stop)) / args.steps
    launches = (tower.launches - launches_before) // args.steps
    final_loss = float(loss.mean().item())
    clocks = sampler.stop() if sampler else None

    # ---- instrumented pass (event pair per launch, everything serialised on the compute stream) ----
    per_kernel_ms, ms_instrumented = arm.instrumented(args.steps)

    # ---- end-to-end arm (public API, host inputs) ----
    names = arm.names

    def e2e_run(steps):
        # the public training call: pipelined H2D of the next batch + async loss read-back
        host_inputs = dict(arm.inputs)
        host_inputs[names.input_batch] = arm.host_x
        return net.fit_batches((host_inputs for _ in range(steps)), global_batch_size=arm.global_batch,
                               data_parallel=dp if dp.active else None)

    def e2e_step():
        host_inputs = dict(arm.inputs)
        host_inputs[names.input_batch] = arm.host_x
        return net.train_on_batch(host_inputs, global_batch_size=arm.global_batch,
                                  data_parallel=dp if dp.active else None)

    if args.e2e_mode == "step":
        for _ in range(max(3, args.warmup // 2)):
            e2e_step()
    else:
        e2e_run(max(3, args.warmup // 2))
    torch.cuda.synchronize()
    dp.barrier()
    t0 = time.perf_counter()
    start.record()
    if args.e2e_mode == "step":
        for _ in range(args.steps):
            e2e_step()
    else:
        e2e_run(args.steps)
    stop.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms_e2e = dp.max_over_ranks(max(start.elapsed_time(stop), wall)) / args.steps
    h2d = arm.host_x.numel() * 4 + arm.inputs[names.label_batch].nbytes + 2 * batch * 4
    d2h = 4

    # ---- the other precision modes on the same workload, the other BASELINE configurations ----
    twins, configs = [], []
    sub_steps, sub_warmup = max(5, min(args.steps, 20)), 3
    if not args.no_twins:
        for twin_dtype in [d for d in ("fp16", "bf16", "bf16x2") if d != dtype]:
            twin = Arm(args.workload, batch, twin_dtype, dp, device)
            twin_ms, _ = twin.timed(sub_steps, sub_warmup)
            twin_kernels, _ = twin.instrumented(5)
            _, conv = twin.conv_summary(twin_kernels, peaks)
            twins.append({"dtype": twin_dtype, "ms_per_step": round(twin_ms, 4),
                          "value": twin.frames_per_step / (twin_ms * 1e-3), "unit": "frames/s",
                          "mma_terms_per_product": 3 if twin_dtype == "bf16x2" else 1, "all_conv": conv,
                          "steps": sub_steps, "warmup": sub_warmup})
            del twin
            torch.cuda.empty_cache()
    if not args.no_configs:
        for name in [w for w in ("small", "long", "ragged") if w != args.workload]:
            # long-form: BASELINE config 5 is a global batch of 128 on 4 / 8 GPUs; 16 per GPU otherwise
            sub_batch = WORKLOADS[name][3]
            if name == "long" and world in (4, 8):
                sub_batch = 128 // world
            sub = Arm(name, sub_batch, WORKLOADS[name][4], dp, device)
            sub_ms, _ = sub.timed(sub_steps, sub_warmup)
            sub_kernels, _ = sub.instrumented(3)
            _, conv = sub.conv_summary(sub_kernels, peaks)
            ctc_ms = per_kernel_ctc(sub_kernels)
            ctc_bytes = ctc_algorithmic_bytes(sub.frames, sub.examples)
            varying = varying_shapes_run(sub, sub_steps) if name == "ragged" else None
            configs.append({"config": config_record(name, sub_batch, world, sub.frames, len(sub.examples[0].label)),
                            "varying_shapes": varying,
                            "dtype": sub.dtype, "ms_per_step": round(sub_ms, 4),
                            "value": sub.frames_per_step / (sub_ms * 1e-3),
                            "value_padded_frames": sub.padded_frames_per_step / (sub_ms * 1e-3), "unit": "frames/s",
                            "all_conv": conv,
                            "ctc": {"ms": round(ctc_ms, 4), "algorithmic_bytes": ctc_bytes,
                                    "frac": round(ctc_bytes / (ctc_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)},
                            "steps": sub_steps, "warmup": sub_warmup})
            del sub
            torch.cuda.empty_cache()

    equivalence = dp_equivalence(dp, device) if (dp.active and not args.no_dp_check) else None

    if dp.active:
        dp.barrier()
        dp.close()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- per-kernel breakdown + roofline of the dominant kernel ----
    flops, all_conv = arm.conv_summary(per_kernel_ms, peaks)
    kernels = []
    for (kind, name), ms in per_kernel_ms.items():
        item = {"kernel": "{}:{}".format(kind, name), "ms": round(ms, 4), "share": round(ms / ms_instrumented, 4)}
        if (kind, name) in flops:
            item["tflops"] = round(flops[(kind, name)] / (ms * 1e-3) / 1e12, 1)
        kernels.append(item)
    kernels.sort(key=lambda k: -k["ms"])
    top = kernels[0]
    top_kind, top_name = top["kernel"].split(":")
    # burst peak unless the timed region is seconds long (then the power cap bites and the sustained figure applies)
    region_s = ms_device * args.steps * 1e-3
    peak_key = "bf16_tflops_sustained" if region_s >= 2.0 else "bf16_tflops"
    peak_tf = peaks[peak_key]
    terms = 3 if dtype == "bf16x2" else 1
    roofline = {
        "kernel": top["kernel"], "bound": "tensor", "achieved": top.get("tflops"), "peak": peak_tf, "unit": "TFLOP/s",
        "frac": round(top["tflops"] / peak_tf, 4) if "tflops" in top else None,
        "traffic": ncu_traffic(top["kernel"], args.workload, dtype),
        "algorithmic_flops": flops.get((top_kind, top_name)),
        "peak_source": "{} of MEASURED_PEAKS.json ({}); timed region {:.2f} s".format(peak_key, peaks["source"], region_s),
        "peak_burst": peaks["bf16_tflops"], "peak_sustained": peaks["bf16_tflops_sustained"],
        "frac_burst": round(top["tflops"] / peaks["bf16_tflops"], 4) if "tflops" in top else None,
        "frac_sustained": round(top["tflops"] / peaks["bf16_tflops_sustained"], 4) if "tflops" in top else None,
        "mma_terms_per_product": terms,
        "timing": "CUDA-event pair around each launch, {} instrumented steps right after the timed region "
                  "({:.3f} ms/step instrumented vs {:.3f} uninstrumented)".format(args.steps, ms_instrumented, ms_device),
        "all_conv": all_conv,
    }
    ctc_ms = per_kernel_ctc(per_kernel_ms)
    if ctc_ms:
        ctc_bytes = ctc_algorithmic_bytes(arm.frames, arm.examples)
        roofline["ctc"] = {"bound": "hbm", "achieved": round(ctc_bytes / (ctc_ms * 1e-3) / 1e9, 1),
                           "peak": peaks["hbm_gbs"], "unit": "GB/s",
                           "frac": round(ctc_bytes / (ctc_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                           "algorithmic_bytes": ctc_bytes, "ms": round(ctc_ms, 4),
                           "traffic": ncu_traffic("ctc:ctc_loss", args.workload, dtype)}
        # the two kernels behind that number: the lattice walk is a chain of T' dependent steps (latency), the
        # gradient kernel streams the alpha / beta rows (bandwidth) — DESIGN.md §4.3
        try:
            lattice_ms = arm.ctc_lattices_only_ms()
        except Exception as error:  # (an extra measurement must never cost the line)
            print("ctc decomposition skipped: {}".format(error), file=sys.stderr)
            lattice_ms = 0.0
        if 0 < lattice_ms < ctc_ms:
            V = len(arm.alphabet) + 1
            gradient_bytes = sum((f // 2) * (2 * (2 * len(e.label) + 1) * 4 + 2 * V * 4)
                                 for f, e in zip(arm.frames, arm.examples))
            gradient_ms = ctc_ms - lattice_ms
            roofline["ctc"]["lattice_kernel"] = {
                "ms": round(lattice_ms, 4), "dependent_steps": max(arm.frames) // 2,
                "ns_per_step": round(lattice_ms * 1e6 / (max(arm.frames) // 2), 1), "bound": "latency",
                "timing": "sl_ctc_loss without the gradient output, 20 back-to-back calls after the timed region"}
            roofline["ctc"]["gradient_kernel"] = {
                "ms": round(gradient_ms, 4), "bound": "hbm", "algorithmic_bytes": gradient_bytes,
                "achieved": round(gradient_bytes / (gradient_ms * 1e-3) / 1e9, 1), "unit": "GB/s",
                "frac": round(gradient_bytes / (gradient_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                "timing": "ctc ms of the instrumented steps minus the lattice kernel's"}

    precision = None if args.no_precision_check else precision_record(dtype, device)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = cpu_reference_rate(args.workload, steps=3, warmup=1, sample_batch=min(16, batch),
                                       budget_seconds=40.0)

    line = {
        "metric": METRIC, "value": arm.frames_per_step / (ms_device * 1e-3), "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_device, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": DATA,
        "config": config_record(args.workload, batch, world, arm.frames, len(arm.examples[0].label)),
        "clocks": clocks,
        "e2e": {"value": arm.frames_per_step / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "host_enqueue_ms_per_step": round(host_enqueue_ms, 3),
        "roofline": roofline,
        "precision": precision,
        "twins": twins,
        "configs": configs,
        "dp_check": equivalence,
        "cpu_baseline": cpu,
        "kernels": kernels[:12],
        "final_loss": final_loss,
    }
    print(json.dumps(line), flush=True)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=40)
    parser.add_argument("--warmup", type=int, default=5)
    parser.add_argument("--impl", default="ours", choices=["ours", "reference"])
    parser.add_argument("--workload", default="full", choices=sorted(WORKLOADS))
    parser.add_argument("--batch-per-gpu", type=int, default=None)
    parser.add_argument("--dtype", default=None, choices=["fp16", "bf16", "bf16x2"])
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--no-twins", action="store_true", help="skip the other precision modes of the workload")
    parser.add_argument("--no-configs", action="store_true", help="skip the short runs of the other configurations")
    parser.add_argument("--no-precision-check", action="store_true")
    parser.add_argument("--no-dp-check", action="store_true")
    parser.add_argument("--overlap-backward", type=int, default=0, help="run wgrad on a side stream (experiment)")
    parser.add_argument("--pipeline-update", type=int, default=1,
                        help="per-bucket all-reduce + Adam on an update stream, overlapped with backward")
    parser.add_argument("--e2e-mode", default="fit", choices=["fit", "step"],
                        help="e2e arm: Wav2Letter.fit_batches (pipelined) or one train_on_batch call per step")
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
