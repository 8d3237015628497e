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


def _surface_worker(rank, world, port, out_dir):
    """The reference-facing data-parallel host logic without a GPU: sharded evaluation (batches dealt
    round-robin, results gathered in order), the shard + global-pad batch assembly of `train`, and the
    stream-ordered all-reduce primitives on their torch.distributed (gloo) fallback."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from speechless_b200 import english_frequent_characters as alphabet
    from speechless_b200.distributed import DataParallel
    from speechless_b200.grapheme_enconding import CtcGraphemeEncoding
    from speechless_b200.net import Wav2Letter
    from speechless_b200.results import ExpectationVsPrediction, ExpectationsVsPredictions
    from speechless_b200.synthetic import synthetic_batch
    dp = DataParallel(backend="gloo")
    assert not dp.owns_communicator  # no CUDA: the exchange goes through torch.distributed

    class Stub:  # the attributes the two public methods touch, without a device
        data_parallel = dp
        grapheme_encoding = CtcGraphemeEncoding(alphabet)
        input_to_prediction_length_ratio = 2
        seen = []

        def test_and_predict_batch_with_log(self, index, batch):
            self.seen.append(index)
            return ExpectationsVsPredictions([ExpectationVsPrediction(predicted=e.label.upper(), expected=e.label,
                                                                      loss=float(index)) for e in batch])
        _input_batch_and_prediction_lengths = Wav2Letter._input_batch_and_prediction_lengths
        _prediction_length_batch = Wav2Letter._prediction_length_batch
        _input_dictionary_for_loss_net = Wav2Letter._input_dictionary_for_loss_net
        _inputs_for_loss_net = Wav2Letter._inputs_for_loss_net

    stub = Stub()
    batches = [synthetic_batch(3, [40 + 2 * s, 31, 25 + s], alphabet, seed=s, label_length=4) for s in range(5)]
    result = Wav2Letter.test_and_predict_batches(stub, batches)
    assert stub.seen == [i for i in range(5) if i % world == rank]          # this rank's share only
    assert [r.loss for b in result.result_batches for r in b.results] == [float(i) for i in range(5) for _ in range(3)]
    assert [r.expected for r in result.results] == [e.label for b in batches for e in b]  # complete and ordered

    generated = list(Wav2Letter._loss_inputs_generator(stub, batches[:2], dp))
    for (inputs, dummy), batch in zip(generated, batches[:2]):
        shard = dp.shard(batch)
        longest = max(e.z_normalized_transposed_spectrogram().shape[0] for e in batch)
        assert inputs["global_batch_size"] == 3 and len(dummy) == len(shard)
        assert inputs[Wav2Letter.InputNames.input_batch].shape == (len(shard), longest, 128)  # GLOBAL pad length
        assert inputs[Wav2Letter.InputNames.prediction_lengths][:, 0].tolist() == [
            e.z_normalized_transposed_spectrogram().shape[0] // 2 for e in shard]

    grads = torch.arange(10, dtype=torch.float32) * (rank + 1)
    dp.allreduce_range(grads, 2, 7)
    want = torch.arange(10, dtype=torch.float32) * (rank + 1)
    want[2:7] = torch.arange(2, 7, dtype=torch.float32) * sum(r + 1 for r in range(world))
    assert torch.equal(grads, want)
    scalar = torch.tensor([float(rank + 1)])
    dp.allreduce_scalar(scalar)
    assert scalar.item() == sum(r + 1 for r in range(world))
    assert dp.gather_objects({"rank": rank}) == [{"rank": r} for r in range(world)]
    dp.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_public_surface_sharding():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_surface_worker, args=(2, port, ""), nprocs=2, join=True)


def test_arena_views_only_grow():
    """Per-shape workspaces are views of arena buffers (engine._Arena): a smaller or equal request reuses the
    buffer, a larger one replaces it and bumps the generation so that stale views get rebuilt."""
    from speechless_b200.engine import _Arena
    arena = _Arena(torch.device("cpu"))
    a = arena.take("act", (2, 100, 64), torch.bfloat16)
    generation = arena.generation
    b = arena.take("act", (2, 90, 64), torch.float16)       # same bytes or fewer: same storage, any 16-bit type
    assert b.data_ptr() == a.data_ptr() and arena.generation == generation and b.shape == (2, 90, 64)
    c = arena.take("act", (2, 105, 64), torch.bfloat16)     # within the 12.5 % headroom: still no allocation
    assert c.data_ptr() == a.data_ptr() and arena.generation == generation
    d = arena.take("act", (4, 200, 64), torch.bfloat16)
    assert d.data_ptr() != a.data_ptr() and arena.generation == generation + 1
    assert arena.take("other", (8,), torch.int32).numel() == 8
