"""N>1 host logic on CPU: world_size-2 `gloo` processes shard a batch by utterance, each computes
its shard's gradient of the GLOBAL mean objective with the oracle, and the SUM all-reduce of
`speechless_b200.distributed.DataParallel` must reproduce the single-process gradient."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent

from speechless_b200.distributed import shard_bounds


def test_shard_bounds_cover_batch():
    for count in (1, 2, 7, 64, 513):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(count, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _case():
    from oracle import keras_tf_oracle as oracle
    rng = np.random.default_rng(21)
    model = oracle.Wav2LetterOracle(6, 5, main_filter_count=4, out_filter_count=6, seed=3)
    for b in model.biases:
        b += 0.05
    x = rng.standard_normal((4, 50, 6))
    labels = np.array([[0, 1, 2], [3, 3, -1], [1, -1, -1], [2, 0, 1]], dtype=np.int32)
    return model, x, labels, [25, 25, 20, 25], [3, 2, 1, 3]


def _flat(dws, dbs):
    return np.concatenate([g.reshape(-1) for g in dws + dbs])


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from speechless_b200.distributed import DataParallel
    dp = DataParallel(backend="gloo", bucket_bytes=256)  # tiny buckets: exercises the bucket loop
    model, x, labels, pred, ll = _case()
    begin, end = shard_bounds(len(x), dp.rank, dp.world_size)
    assert dp.shard(list(range(len(x)))) == list(range(begin, end))
    # shard padded to the GLOBAL max T (all utterances here share T) ; gradient of (1/B_global) * sum loss
    losses, _, _, dws, dbs = model.loss_and_gradients(x[begin:end], labels[begin:end], pred[begin:end], ll[begin:end])
    shard_b = end - begin
    grads = torch.from_numpy(_flat(dws, dbs) * shard_b / len(x))  # oracle returns d(mean over shard)
    loss_sum = torch.tensor([losses.sum()])
    dp.allreduce(grads, loss_sum)
    assert dp.max_over_ranks(float(rank)) == world - 1
    dp.barrier()
    np.save(os.path.join(out_dir, "grads{}.npy".format(rank)), grads.numpy())
    np.save(os.path.join(out_dir, "loss{}.npy".format(rank)), loss_sum.numpy())
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_allreduce_equals_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    model, x, labels, pred, ll = _case()
    losses, _, _, dws, dbs = model.loss_and_gradients(x, labels, pred, ll)
    want = _flat(dws, dbs)
    for rank in range(2):
        got = np.load(tmp_path / "grads{}.npy".format(rank))
        assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max())
        assert abs(np.load(tmp_path / "loss{}.npy".format(rank))[0] - losses.sum()) < 1e-9
