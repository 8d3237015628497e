#!/usr/bin/env python
"""Data-parallel equivalence check (run under torchrun, one rank per GPU):
the SUM all-reduce of per-shard gradients of the GLOBAL mean CTC objective must equal the
single-GPU gradient of the full batch, and replicas must stay identical after Adam steps.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dp_check.py
"""
import json
import sys
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
    examples = synthetic_batch(global_batch, [300 - 7 * i for i in range(global_batch)], alphabet, seed=5,
                               label_length=20)
    kwargs = dict(main_filter_count=128, out_filter_count=256, seed=3, device=device, compute_dtype="bf16x2")
    names = Wav2Letter.InputNames

    def step_inputs(net, batch, pad_to):
        inputs, _ = net._inputs_for_loss_net(batch)
        x = inputs[names.input_batch]
        if x.shape[1] < pad_to:  # every shard is padded to the GLOBAL max T
            padded = np.zeros((x.shape[0], pad_to, x.shape[2]), dtype=x.dtype)
            padded[:, :x.shape[1]] = x
            inputs[names.input_batch] = padded
        return inputs

    max_t = max(e.z_normalized_transposed_spectrogram().shape[0] for e in examples)
    # --- data parallel: each rank trains on its shard
    net = Wav2Letter(128, alphabet, **kwargs)
    shard = dp.shard(examples)
    losses = []
    for _ in range(3):
        losses.append(net.train_on_batch(step_inputs(net, shard, max_t), global_batch_size=global_batch,
                                         allreduce=dp.allreduce))
    grads_dp = net.tower.grads.clone()
    params_dp = net.tower.params.clone()
    loss_dp = torch.tensor(losses, device=device, dtype=torch.float64)  # already the global batch mean

    # --- replicas identical?
    reference = params_dp.clone()
    dist.broadcast(reference, src=0)
    replicas_identical = bool(torch.equal(reference, params_dp))

    # --- single GPU on the full batch (every rank computes it; compare on each)
    single = Wav2Letter(128, alphabet, **kwargs)
    single_losses = [single.train_on_batch(step_inputs(single, examples, max_t)) for _ in range(3)]
    grads_single = single.tower.grads
    scale = float(grads_single.abs().max())
    grad_err = float((grads_dp - grads_single).abs().max()) / scale
    param_err = float((params_dp - single.tower.params).abs().max())
    loss_err = float(np.abs(loss_dp.cpu().numpy() / np.array(single_losses) - 1).max())
    ok = replicas_identical and grad_err < 2e-3 and loss_err < 1e-5 and param_err < 5e-4
    gathered = [None] * dp.world_size
    dist.all_gather_object(gathered, dict(rank=dp.rank, ok=ok, replicas_identical=replicas_identical,
                                          grad_rel_err=grad_err, loss_rel_err=loss_err, param_abs_err=param_err))
    if dp.rank == 0:
        print(json.dumps({"world_size": dp.world_size, "ok": all(g["ok"] for g in gathered), "ranks": gathered}))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
