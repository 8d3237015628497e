"""Data parallelism by utterance: one process per GPU and ONE exchange step — the sum of
the Conv1D gradients (and of the per-shard loss sums) over NVLink 5 / NVSwitch.

The reference is single-process (no collective anywhere); the mean-over-batch
objective of net.py:389 makes the path data parallel: every rank scales its CTC gradient by
1/B_global, so a plain SUM all-reduce gives every replica the gradient of the single-GPU
step on the full batch, and identical Adam updates keep replicas bit-identical.

Two layers:
  * `torch.distributed` is the plumbing: rendezvous, barriers, scalar max over ranks, and the
    side channel that hands rank 0's communicator id to the other ranks.
  * The data plane is the library's own communicator (`sl_comm_init_rank` / `sl_allreduce_sum`,
    include/speechless_b200.h): NCCL created with a bounded CTA count (`max_ctas`, default 16), because
    the collective runs next to persistent one-CTA-per-SM tensor-core kernels.  Without CUDA (the
    world-size-2 gloo tests) the all-reduce goes through `torch.distributed` instead.
"""
import ctypes
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
    def __init__(self, backend: Optional[str] = None, bucket_bytes: int = 32 << 20, max_ctas: Optional[int] = None,
                 limited_launches: Optional[int] = None, own_communicator: Optional[bool] = None):
        self.rank, self.world_size, self.local_rank = env_world()
        self.bucket_bytes = bucket_bytes
        # CTAs the all-reduce kernels may occupy, and how many of the backward launches following the
        # start of a bucket's all-reduce run on 148 - max_ctas CTAs (ConvTower.backward_and_update)
        self.max_ctas = int(os.environ.get("SL_COMM_MAX_CTAS", "16")) if max_ctas is None else max_ctas
        self.limited_launches = int(os.environ.get("SL_COMM_LIMITED_LAUNCHES", "2")) \
            if limited_launches is None else limited_launches
        self._pending = []
        self._comm = None  # opaque handle of the C-ABI communicator
        self._comm_stream = None
        self._lib = None
        if self.world_size > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world_size)
        if own_communicator is None:
            own_communicator = os.environ.get("SL_OWN_COMM", "1") != "0"
        if self.active and torch.cuda.is_available() and dist.get_backend() == "nccl" and own_communicator:
            self._init_communicator()

    # ------------------------------------------------------------------ the library's communicator
    def _init_communicator(self) -> None:
        from speechless_b200 import _lib
        self._lib = _lib.load()
        torch.cuda.set_device(self.local_rank)
        ident = ctypes.create_string_buffer(_lib.COMM_ID_BYTES)
        if self.rank == 0:
            _lib.check(self._lib.sl_comm_unique_id(ident))
        box = [ident.raw]
        dist.broadcast_object_list(box, src=0)  # the side channel for the 128-byte id
        ident = ctypes.create_string_buffer(box[0], _lib.COMM_ID_BYTES)
        handle = ctypes.c_void_p()
        _lib.check(self._lib.sl_comm_init_rank(ctypes.byref(handle), ident, self.world_size, self.rank, self.max_ctas))
        self._comm = handle
        if self._lib.sl_comm_size(self._comm) != self.world_size:
            raise RuntimeError("communicator has {} ranks, expected {}".format(
                self._lib.sl_comm_size(self._comm), self.world_size))
        self._comm_stream = torch.cuda.Stream()

    @property
    def owns_communicator(self) -> bool:
        return self._comm is not None

    def close(self) -> None:
        if self._comm is not None:
            torch.cuda.synchronize()
            self._lib.sl_comm_destroy(self._comm)
            self._comm = None

    @property
    def active(self) -> bool:
        return self.world_size > 1

    def shard(self, batch: Sequence) -> List:
        begin, end = shard_bounds(len(batch), self.rank, self.world_size)
        return list(batch[begin:end])

    # ------------------------------------------------------------------ stream-ordered primitives
    def allreduce_range(self, grads: torch.Tensor, begin: int, end: int, stream=None) -> None:
        """In-place SUM all-reduce of grads[begin:end] (fp32), ordered on `stream` (default: the current one)."""
        if not self.active:
            return
        flat = grads.view(-1)
        if self._comm is None:
            work = dist.all_reduce(flat[begin:end], op=dist.ReduceOp.SUM, async_op=True)
            work.wait()
            return
        from speechless_b200 import _lib
        s = torch.cuda.current_stream() if stream is None else stream
        _lib.check(self._lib.sl_allreduce_sum(self._comm, flat.data_ptr() + begin * flat.element_size(), end - begin,
                                              s.cuda_stream))

    def allreduce_scalar(self, value: torch.Tensor, stream=None, after=None) -> None:
        """SUM all-reduce of a one-element fp32 tensor; `after`: a stream whose enqueued work produces it."""
        if not self.active:
            return
        if self._comm is None:
            dist.all_reduce(value, op=dist.ReduceOp.SUM)
            return
        from speechless_b200 import _lib
        s = torch.cuda.current_stream() if stream is None else stream
        if after is not None and after is not s:
            s.wait_event(after.record_event())
        _lib.check(self._lib.sl_allreduce_sum(self._comm, value.data_ptr(), value.numel(), s.cuda_stream))

    # ------------------------------------------------------------------ whole-buffer hooks (non-overlapped path)
    def allreduce(self, grads: torch.Tensor, loss_sum: Optional[torch.Tensor] = None) -> None:
        """SUM over ranks, in place, of the whole flat gradient buffer (and the loss sum) on the current
        stream — the plain hook `Wav2Letter.train_on_batch(allreduce=...)` takes."""
        if not self.active:
            return
        flat = grads.view(-1)
        if self._comm is not None:
            self.allreduce_range(flat, 0, flat.numel())
            if loss_sum is not None:
                self.allreduce_scalar(loss_sum)
            return
        step = max(1, self.bucket_bytes // flat.element_size())
        handles = [dist.all_reduce(flat[start:start + step], op=dist.ReduceOp.SUM, async_op=True)
                   for start in range(0, flat.numel(), step)]
        if loss_sum is not None:
            handles.append(dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM, async_op=True))
        for handle in handles:
            handle.wait()

    def allreduce_bucket_async(self, grads: torch.Tensor, begin: int, end: int) -> None:
        """Start the SUM all-reduce of grads[begin:end] behind the work enqueued so far (called from
        `ConvTower.backward` as each bucket's weight gradients are enqueued); `finish()` makes the
        compute stream wait."""
        if not self.active:
            return
        if self._comm is None:
            self._pending.append(dist.all_reduce(grads.view(-1)[begin:end], op=dist.ReduceOp.SUM, async_op=True))
            return
        self._comm_stream.wait_event(torch.cuda.current_stream().record_event())
        self.allreduce_range(grads, begin, end, stream=self._comm_stream)

    def finish(self, loss_sum: Optional[torch.Tensor] = None) -> None:
        if not self.active:
            return
        if self._comm is not None:
            if loss_sum is not None:
                self.allreduce_scalar(loss_sum, stream=self._comm_stream, after=torch.cuda.current_stream())
            torch.cuda.current_stream().wait_stream(self._comm_stream)
            return
        if loss_sum is not None:
            self._pending.append(dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM, async_op=True))
        for handle in self._pending:
            handle.wait()
        self._pending = []

    # ------------------------------------------------------------------ plumbing
    def max_over_ranks(self, value: float) -> float:
        if not self.active:
            return value
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_objects(self, obj) -> List:
        """Every rank's `obj` on every rank (host side: strings, losses — evaluation needs no collective
        on the data path)."""
        if not self.active:
            return [obj]
        gathered = [None] * self.world_size
        dist.all_gather_object(gathered, obj)
        return gathered

    def barrier(self) -> None:
        if self.active:
            dist.barrier()
