#!/usr/bin/env python
"""bench.py — spectrogram frames/sec of the wav2letter hot path (fwd + CTC + bwd + all-reduce + Adam).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload full|small|long] [--batch-per-gpu B] [--dtype bf16|bf16x2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  Metric / workloads: BASELINE.json.  A "step" is one
training step on one synthetic batch of B utterances per GPU (weak scaling):
  value  — frames/s with the batch already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e    — frames/s through the public `Wav2Letter.train_on_batch` call with HOST inputs
           (pinned host -> device copy of the batch and device -> host read of the loss inside
           the timed region)
  roofline — the dominant kernel (largest share of the step) against MEASURED_PEAKS.json
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
    "full": (250, 2000, 10.0, 64, "bf16", "wav2letter-full (250/2000 filters) batch=64/GPU, 10 s utterances"),
    "small": (256, 256, 10.0, 32, "bf16x2", "wav2letter-small (11 Conv1D, 256ch, mel-128) batch=32, 10 s, fp32-parity"),
    "long": (250, 2000, 60.0, 16, "bf16", "long-form 60 s utterances (T=7501, T'=3751), CTC stress"),
}


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
    the committed `ncu --set full` capture summary (profiles/ncu_traffic.json), or None."""
    path = ROOT / "profiles" / "ncu_traffic.json"
    if not path.exists():
        return None
    table = json.loads(path.read_text())
    entry = table.get("{}|{}|{}".format(workload, dtype, kernel_key))
    return entry["dram_bytes"] if entry else None


def make_host_batch(args, rank):
    import numpy as np
    from speechless_b200 import english_frequent_characters as alphabet
    from speechless_b200.synthetic import frames_for_seconds, synthetic_batch
    main, out, seconds, default_batch, default_dtype, _ = WORKLOADS[args.workload]
    frames = frames_for_seconds(seconds)
    batch = args.batch_per_gpu or default_batch
    return synthetic_batch(batch, frames, alphabet, seed=1234 + rank), frames, batch, alphabet


def cpu_reference_rate(args, steps, warmup, sample_batch=2, budget_seconds=25.0):
    """frames/s of the torch-CPU restatement (oracle/torch_cpu.py) on a bounded sample of the workload."""
    import numpy as np
    import torch
    from oracle.torch_cpu import TorchCpuWav2Letter
    from speechless_b200 import english_frequent_characters as alphabet
    from speechless_b200.grapheme_enconding import CtcGraphemeEncoding
    from speechless_b200.synthetic import frames_for_seconds, synthetic_batch
    main, out, seconds, _, _, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    frames = frames_for_seconds(seconds)
    batch = synthetic_batch(sample_batch, frames, alphabet, seed=1234)
    encoding = CtcGraphemeEncoding(alphabet)
    x = np.stack([e.z_normalized_transposed_spectrogram() for e in batch]).astype(np.float32)
    labels = encoding.encode_label_batch([e.label for e in batch])
    pred_len = [frames // 2] * sample_batch
    label_len = [len(e.label) for e in batch]
    model = TorchCpuWav2Letter(128, len(alphabet) + 1, main, out, seed=0)
    for _ in range(warmup):
        model.train_step(x, labels, pred_len, label_len)
    times = []
    started = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        model.train_step(x, labels, pred_len, label_len)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - started > budget_seconds:
            break
    mean = sum(times) / len(times)
    return {"value": sample_batch * frames / mean, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": "torch-CPU fp32 restatement of the Keras/TF path (Keras/TF not installable offline); "
                      "{} utterances x {} frames per step, {} timed step(s), {:.2f} s/step".format(
                sample_batch, frames, len(times), mean)}, mean, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    main, out, seconds, default_batch, default_dtype, config_name = WORKLOADS[args.workload]
    steps = min(args.steps, 20)
    baseline, mean, timed = cpu_reference_rate(args, steps=steps, warmup=min(args.warmup, 1), budget_seconds=120.0)
    line = {
        "impl": "reference", "metric": "spectrogram frames/sec, fwd+bwd+CTC+Adam training step",
        "value": baseline["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": timed,
        "warmup": min(args.warmup, 1), "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic N(0,1) mel-128 spectrograms, uniform random labels",
        "config": {"workload": config_name, "note": "bounded CPU sample of the same workload"},
        "cpu_baseline": baseline,
        "e2e": {"value": baseline["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from speechless_b200.distributed import DataParallel
    from speechless_b200.net import Wav2Letter

    # NCCL prints its version banner on stdout when NCCL_DEBUG=VERSION/INFO; stdout must carry
    # exactly one JSON line
    os.environ["NCCL_DEBUG"] = os.environ.get("SL_NCCL_DEBUG", "WARN")
    dp = DataParallel()
    rank, world = dp.rank, dp.world_size
    if world != args.gpus and rank == 0:
        print("warning: --gpus {} but WORLD_SIZE {}".format(args.gpus, world), file=sys.stderr)
    torch.cuda.set_device(dp.local_rank)
    device = torch.device("cuda", dp.local_rank)

    main, out, seconds, default_batch, default_dtype, config_name = WORKLOADS[args.workload]
    dtype = args.dtype or default_dtype
    examples, frames, batch, alphabet = make_host_batch(args, rank)
    global_batch = batch * world

    net = Wav2Letter(128, alphabet, main_filter_count=main, out_filter_count=out, compute_dtype=dtype,
                     device=device, seed=0)
    tower = net.tower
    tower.overlap_backward = bool(args.overlap_backward)
    inputs, _ = net._inputs_for_loss_net(examples)
    names = Wav2Letter.InputNames
    host_x = torch.from_numpy(inputs[names.input_batch]).pin_memory()
    allreduce = dp.allreduce if dp.active else None

    def device_step(ws):
        tower.upload(ws.x_f32)  # re-pack from the HBM-resident fp32 batch
        tower.forward(ws)
        loss = tower.ctc(ws, want_grad=True, grad_scale=1.0 / global_batch)
        if dp.active and args.overlap_allreduce:
            tower.backward(ws, on_bucket_ready=lambda begin, end: dp.allreduce_bucket_async(tower.grads, begin, end))
            dp.finish()
        else:
            tower.backward(ws)
            if allreduce is not None:
                allreduce(tower.grads, None)
        net.optimizer.iterations += 1
        tower.adam_step(net.optimizer.lr, net.optimizer.beta_1, net.optimizer.beta_2, net.optimizer.epsilon,
                        net.optimizer.iterations)
        return loss

    def e2e_step():
        host_inputs = dict(inputs)
        host_inputs[names.input_batch] = host_x
        if dp.active and args.overlap_allreduce:
            return net.train_on_batch(host_inputs, global_batch_size=global_batch, data_parallel=dp)
        return net.train_on_batch(host_inputs, global_batch_size=global_batch, allreduce=allreduce)

    # ---- device-resident arm ----
    ws = tower.upload(host_x)
    tower.set_labels(ws, inputs[names.label_batch], inputs[names.prediction_lengths], inputs[names.label_lengths])
    sampler = ClockSampler(dp.local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        device_step(ws)
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
        loss = device_step(ws)
    stop.record()
    host_enqueue_ms = (time.perf_counter() - host_t0) * 1e3 / args.steps
    torch.cuda.synchronize()
    dp.barrier()
    torch.cuda.synchronize()
    ms_device = dp.max_over_ranks(start.elapsed_time(stop)) / args.steps
    launches = (tower.launches - launches_before) // args.steps
    final_loss = float(loss.mean().item())

    # ---- instrumented pass: the same K steps with a CUDA-event pair around every launch (the
    # ~100 extra timestamp operations per step cost a few percent, so they stay out of `value`)
    tower.profile = []
    start.record()
    for _ in range(args.steps):
        device_step(ws)
    stop.record()
    torch.cuda.synchronize()
    dp.barrier()
    ms_instrumented = start.elapsed_time(stop) / args.steps
    clocks = sampler.stop() if sampler else None
    profile, tower.profile = tower.profile, None

    # ---- end-to-end arm (public API, host inputs) ----
    def e2e_run(steps):
        # the public training call: pipelined H2D of the next batch + async loss read-back
        host_inputs = dict(inputs)
        host_inputs[names.input_batch] = host_x
        return net.fit_batches((host_inputs for _ in range(steps)), global_batch_size=global_batch,
                               data_parallel=dp if (dp.active and args.overlap_allreduce) else None)

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
    h2d = host_x.numel() * 4 + inputs[names.label_batch].nbytes + 2 * batch * 4
    d2h = 4

    if dp.active:
        dp.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- per-kernel breakdown + roofline of the dominant kernel ----
    peaks = measured_peaks()
    flops = conv_flops(tower.layers, ws.t_out, batch, tower.first_trainable())
    agg = {}
    for kind, name, ev0, ev1 in profile:
        entry = agg.setdefault((kind, name), [0.0, 0])
        entry[0] += ev0.elapsed_time(ev1)
        entry[1] += 1
    kernels = []
    for (kind, name), (total_ms, count) in agg.items():
        ms = total_ms / count
        item = {"kernel": "{}:{}".format(kind, name), "ms": round(ms, 4), "share": round(ms / ms_instrumented, 4)}
        if (kind, name) in flops:
            item["tflops"] = round(flops[(kind, name)] / (ms * 1e-3) / 1e12, 1)
        kernels.append(item)
    kernels.sort(key=lambda k: -k["ms"])
    conv_ms = sum(k["ms"] for k in kernels if "tflops" in k)
    conv_flop_total = sum(flops.values())
    top = kernels[0]
    top_kind, top_name = top["kernel"].split(":")
    peak_tf = peaks["bf16_tflops_sustained"]  # timed inside a long step -> sustained figure
    terms = 3 if dtype == "bf16x2" else 1
    roofline = {
        "kernel": top["kernel"], "bound": "tensor", "achieved": top.get("tflops"), "peak": peak_tf, "unit": "TFLOP/s",
        "frac": round(top["tflops"] / peak_tf, 4) if "tflops" in top else None,
        "traffic": ncu_traffic(top["kernel"], args.workload, dtype),
        "algorithmic_flops": flops.get((top_kind, top_name)),
        "peak_source": "{} ({})".format("bf16_tflops_sustained of MEASURED_PEAKS.json", peaks["source"]),
        # for orientation: the burst figure (a kernel timed alone) and the fraction against it
        "peak_burst": peaks["bf16_tflops"],
        "frac_burst": round(top["tflops"] / peaks["bf16_tflops"], 4) if "tflops" in top else None,
        "mma_terms_per_product": terms,
        "timing": "CUDA-event pair around each launch, {} instrumented steps right after the timed region "
                  "({:.3f} ms/step instrumented vs {:.3f} uninstrumented)".format(args.steps, ms_instrumented, ms_device),
        "all_conv": {"achieved": round(conv_flop_total / (conv_ms * 1e-3) / 1e12, 1),
                     "frac": round(conv_flop_total / (conv_ms * 1e-3) / 1e12 / peak_tf, 4),
                     "ms": round(conv_ms, 3), "flops_per_step": conv_flop_total},
    }
    ctc = [k for k in kernels if k["kernel"].startswith("ctc")]
    if ctc:
        L = len(examples[0].label)
        S = 2 * L + 1
        P = frames // 2
        V = len(alphabet) + 1
        ctc_bytes = batch * (2 * P * V * 4 + 2 * P * S * 4 + L * 4)
        roofline["ctc"] = {"bound": "hbm", "achieved": round(ctc_bytes / (ctc[0]["ms"] * 1e-3) / 1e9, 1),
                           "peak": peaks["hbm_gbs"], "unit": "GB/s",
                           "frac": round(ctc_bytes / (ctc[0]["ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                           "algorithmic_bytes": ctc_bytes, "ms": ctc[0]["ms"]}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = cpu_reference_rate(args, steps=5, warmup=1)

    frames_per_step = global_batch * frames
    line = {
        "metric": "spectrogram frames/sec, fwd+bwd+CTC+Adam training step",
        "value": frames_per_step / (ms_device * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_device, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dtype,
        "data": "synthetic N(0,1) mel-128 spectrograms, uniform random labels, glorot-uniform weights",
        "config": {"workload": config_name, "global_batch": global_batch, "batch_per_gpu": batch,
                   "frames_per_utterance": frames, "label_length": len(examples[0].label),
                   "parallelism": "dp{}".format(world),
                   "l2": "per-step working set (~1 GB of activations) exceeds the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": frames_per_step / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "host_enqueue_ms_per_step": round(host_enqueue_ms, 3),
        "roofline": roofline,
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
    parser.add_argument("--dtype", default=None, choices=["bf16", "bf16x2", "fp16"])
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--overlap-backward", type=int, default=0, help="run wgrad on a side stream (experiment)")
    parser.add_argument("--e2e-mode", default="fit", choices=["fit", "step"],
                        help="e2e arm: Wav2Letter.fit_batches (pipelined) or one train_on_batch call per step")
    parser.add_argument("--overlap-allreduce", type=int, default=1,
                        help="start each gradient bucket's all-reduce as soon as backward produced it")
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
