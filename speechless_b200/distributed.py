"""Data parallelism by utterance: one process per GPU, `torch.distributed`
(NCCL over NVLink 5 / NVSwitch) for the single exchange step of the path — the sum of the
Conv1D gradients (and of the per-shard loss sums).

The reference is single-process (no collective anywhere); the mean-over-batch
objective of net.py:389 makes the path data parallel: every rank scales its CTC gradient by
1/B_global, so a plain SUM all-reduce gives every replica the gradient of the single-GPU
step on the full batch, and identical Adam updates keep replicas bit-identical.
"""
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_bounds(count: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split of `count` utterances; the first `count % world_size` ranks get one more."""
    base, extra = divmod(count, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


class DataParallel:
    def __init__(self, backend: Optional[str] = None, bucket_bytes: int = 32 << 20):
        self.rank, self.world_size, self.local_rank = env_world()
        self.bucket_bytes = bucket_bytes
        self._pending = []
        if self.world_size > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world_size)

    @property
    def active(self) -> bool:
        return self.world_size > 1

    def shard(self, batch: Sequence) -> List:
        begin, end = shard_bounds(len(batch), self.rank, self.world_size)
        return list(batch[begin:end])

    def allreduce(self, grads: torch.Tensor, loss_sum: Optional[torch.Tensor] = None) -> None:
        """SUM over ranks, in place.  The flat gradient buffer goes out in `bucket_bytes` pieces so
        NCCL pipelines them on its own stream; NVSwitch (NVLS) reduces in-switch when available."""
        if not self.active:
            return
        flat = grads.view(-1)
        step = max(1, self.bucket_bytes // flat.element_size())
        handles = [dist.all_reduce(flat[start:start + step], op=dist.ReduceOp.SUM, async_op=True)
                   for start in range(0, flat.numel(), step)]
        if loss_sum is not None:
            handles.append(dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM, async_op=True))
        for handle in handles:
            handle.wait()

    def allreduce_bucket_async(self, grads: torch.Tensor, begin: int, end: int) -> None:
        """Start the SUM all-reduce of grads[begin:end] (called from `ConvTower.backward` as each
        bucket's weight gradients are enqueued); `finish()` makes the compute stream wait."""
        if self.active:
            self._pending.append(dist.all_reduce(grads.view(-1)[begin:end], op=dist.ReduceOp.SUM, async_op=True))

    def finish(self, loss_sum: Optional[torch.Tensor] = None) -> None:
        if not self.active:
            return
        if loss_sum is not None:
            self._pending.append(dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM, async_op=True))
        for handle in self._pending:
            handle.wait()
        self._pending = []

    def max_over_ranks(self, value: float) -> float:
        if not self.active:
            return value
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self) -> None:
        if self.active:
            dist.barrier()
