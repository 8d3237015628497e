#!/usr/bin/env python
"""Data-parallel equivalence check (run under torchrun, one rank per GPU): N GPUs must take exactly the
optimisation steps one GPU takes on the full batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dp_check.py

Three arms against a single-GPU model trained on the full (ragged) batches:
  hook      `train_on_batch(allreduce=dp.allreduce)`: whole-buffer all-reduce after backward
  pipelined `train_on_batch(data_parallel=dp)`: per-bucket all-reduce + Adam on the update stream,
            persistent grids narrowed while a bucket is in flight (`ConvTower.backward_and_update`)
  surface   `Wav2Letter(..., data_parallel=dp).train(...)` + sharded `test_and_predict_batches` — the
            reference's public calls (net.py:521-556); every rank is handed the same batch iterable
Every shard is padded to the global longest utterance (padding is unmasked).
Prints one JSON line (rank 0) and exits non-zero on mismatch.
"""
import json
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np
import torch
import torch.distributed as dist

from speechless_b200 import english_frequent_characters as alphabet
from speechless_b200.distributed import DataParallel
from speechless_b200.net import Wav2Letter
from speechless_b200.synthetic import synthetic_batch


def main():
    dp = DataParallel()
    torch.cuda.set_device(dp.local_rank)
    device = torch.device("cuda", dp.local_rank)
    global_batch = 4 * dp.world_size
    steps = 3
    batches = [synthetic_batch(global_batch, [300 - 7 * i - 3 * s for i in range(global_batch)], alphabet, seed=5 + s,
                               label_length=20) for s in range(steps)]
    kwargs = dict(main_filter_count=128, out_filter_count=256, seed=3, device=device, compute_dtype="bf16x2")
    names = Wav2Letter.InputNames

    def shard_inputs(net, batch):
        longest = max(e.z_normalized_transposed_spectrogram().shape[0] for e in batch)
        return net._input_dictionary_for_loss_net(dp.shard(batch), pad_to_length=longest)

    # --- single GPU on the full batches (every rank computes it; compare on each)
    single = Wav2Letter(128, alphabet, **kwargs)
    single.tower.retain_gradients = True
    single_losses = [single.train_on_batch(single._inputs_for_loss_net(batch)[0]) for batch in batches]
    grads_single = single.tower.grads.clone()
    params_single = single.tower.params.clone()
    scale = float(grads_single.abs().max())

    report = {}
    ok = True
    for arm in ("hook", "pipelined"):
        net = Wav2Letter(128, alphabet, **kwargs)
        net.tower.retain_gradients = True  # (the pipelined step otherwise clears each bucket after its update)
        losses = []
        for batch in batches:
            inputs = shard_inputs(net, batch)
            if arm == "hook":
                losses.append(net.train_on_batch(inputs, global_batch_size=global_batch, allreduce=dp.allreduce))
            else:
                losses.append(net.train_on_batch(inputs, global_batch_size=global_batch, data_parallel=dp))
        torch.cuda.synchronize()
        reference = net.tower.params.clone()
        dist.broadcast(reference, src=0)
        entry = dict(
            replicas_identical=bool(torch.equal(reference, net.tower.params)),
            grad_rel_err=float((net.tower.grads - grads_single).abs().max()) / scale,
            param_abs_err=float((net.tower.params - params_single).abs().max()),
            loss_rel_err=float(np.abs(np.array(losses) / np.array(single_losses) - 1).max()))
        entry["ok"] = entry["replicas_identical"] and entry["grad_rel_err"] < 2e-3 and entry["loss_rel_err"] < 1e-5 \
            and entry["param_abs_err"] < 5e-4
        ok = ok and entry["ok"]
        report[arm] = entry

    # --- the public surface: train() + sharded evaluation
    with tempfile.TemporaryDirectory() as tmp:
        surface = Wav2Letter(128, alphabet, data_parallel=dp, **{k: v for k, v in kwargs.items() if k != "device"})
        surface.train(iter(batches), preview_labeled_spectrogram_batch=batches[0][:2], tensor_board_log_directory=None,
                      net_directory=Path(tmp) / "nets", batches_per_epoch=steps, epochs=1)
        evaluation = surface.test_and_predict_batches(batches)
    torch.cuda.synchronize()
    want = single.test_and_predict_batches(batches)
    got_losses = np.array([r.loss for b in evaluation.result_batches for r in b.results])
    want_losses = np.array([r.loss for b in want.result_batches for r in b.results])
    entry = dict(param_abs_err=float((surface.tower.params - params_single).abs().max()),
                 evaluation_count=int(len(got_losses)),
                 evaluation_loss_rel_err=float(np.abs(got_losses / want_losses - 1).max()),
                 predictions_equal=[r.predicted for b in evaluation.result_batches for r in b.results] ==
                                   [r.predicted for b in want.result_batches for r in b.results],
                 own_communicator=dp.owns_communicator, comm_max_ctas=dp.max_ctas)
    entry["ok"] = entry["param_abs_err"] < 5e-4 and entry["evaluation_loss_rel_err"] < 1e-3 and \
        entry["evaluation_count"] == steps * global_batch
    ok = ok and entry["ok"]
    report["surface"] = entry

    gathered = dp.gather_objects(dict(rank=dp.rank, ok=ok, **report))
    if dp.rank == 0:
        print(json.dumps({"world_size": dp.world_size, "ok": all(g["ok"] for g in gathered), "ranks": gathered}))
    dp.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
