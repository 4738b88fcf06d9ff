"""Batch sharding of the hot path across the GPUs of one box.

Every token / latent is independent given the replicated codebook and entropy parameters, so the
path shards by BATCH with no data-path collective: rank r quantizes / rates images
``shard_bounds(B, r, world)``.  The reference is single-GPU only (README.md:65;
src/trainer/base_trainer.py:158 "TODO: gather loss_dict from all ranks"), so this file is new
host-side plumbing, not a port.  The one real exchange step is in training (config 5): the
gradients of whatever parameters a stage leaves trainable are summed over ranks once per step --
``allreduce_gradients`` does that in flat buckets (NCCL over NVLink on GPUs, gloo in the CPU tests).
One process per GPU, launched by torchrun; rendezvous on 127.0.0.1.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of the batch for `rank`; the first ``batch % world`` ranks get one extra
    image, ranks beyond the batch get an empty slice."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's images of a batch-first tensor (a view, no copy)."""
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world: int | None = None,
                        bucket_bytes: int = 64 << 20, group=None, average: bool = True) -> int:
    """Sum (and by default average) ``p.grad`` over ranks, in flat buckets of at most `bucket_bytes`.

    Parameters without a gradient on this rank contribute zeros (the collective must be issued by every
    rank with the same bucket layout, which depends only on the parameter list).  Returns the number of
    collectives issued.  With NCCL the buckets are device tensors and the all-reduce runs over
    NVLink/NVSwitch; bucket size trades launch latency against overlap, not link count.
    """
    if not dist.is_available() or not dist.is_initialized():
        return 0
    world = world or dist.get_world_size(group)
    if world == 1:
        return 0
    plist: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
    n_coll = 0
    bucket: List[torch.nn.Parameter] = []
    size = 0

    def flush():
        nonlocal bucket, size, n_coll
        if not bucket:
            return
        dev, dt = bucket[0].device, bucket[0].dtype
        flat = torch.zeros(sum(p.numel() for p in bucket), device=dev, dtype=dt)
        off = 0
        for p in bucket:
            if p.grad is not None:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
            off += p.numel()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(world)
        off = 0
        for p in bucket:
            g = flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += p.numel()
        n_coll += 1
        bucket, size = [], 0

    for p in plist:
        nbytes = p.numel() * p.element_size()
        if bucket and (size + nbytes > bucket_bytes or p.dtype != bucket[0].dtype or p.device != bucket[0].device):
            flush()
        bucket.append(p)
        size += nbytes
    flush()
    return n_coll


class GradBucket:
    """A persistent flat FP32 gradient buffer with named views, all-reduced as ONE collective.

    SURVEY 8(e): the kernels that produce gradients write straight into the bucket (``dcvic_vq_backward`` takes any
    ``dE`` pointer, so the codebook gradient's scatter-add lands in ``view("codebook")``; the entropy-parameter
    gradients are copied into their views), and the all-reduce is launched on a SIDE stream as soon as the producing
    kernels are enqueued, so it overlaps whatever the compute stream does next (``dz`` of the next micro-batch, the
    decoder's backward).  ``wait()`` makes the compute stream wait for the collective before the views are read."""

    def __init__(self, device, fields, group=None, average: bool = True):
        self.device = torch.device(device)
        self.group, self.average = group, average
        self.offsets, off = {}, 0
        for name, shape in fields:
            n = 1
            for d in shape:
                n *= int(d)
            self.offsets[name] = (off, n, tuple(int(d) for d in shape))
            off += (n + 3) // 4 * 4                      # 16-byte aligned views (128-bit kernels write them)
        self.flat = torch.zeros(max(off, 1), dtype=torch.float32, device=self.device)
        self._side = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._done = None
        self._work = None

    def view(self, name: str) -> torch.Tensor:
        off, n, shape = self.offsets[name]
        return self.flat[off:off + n].view(shape)

    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def allreduce_async(self) -> None:
        """Enqueue the all-reduce behind everything already enqueued on the current stream."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        if self._side is None:                            # CPU tensors (gloo in the tests): synchronous
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                self.flat.div_(world)
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                self.flat.div_(world)
            self._done = torch.cuda.Event()
            self._done.record(self._side)

    def wait(self) -> None:
        if self._done is not None:
            torch.cuda.current_stream(self.device).wait_event(self._done)
            self._done = None


def max_over_ranks(value: float, device: torch.device | str = "cpu", group=None) -> float:
    """Device-timed durations are reported as the maximum over ranks."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t)


def sum_over_ranks(values: Sequence[float], device: torch.device | str = "cpu", group=None) -> List[float]:
    """Per-rank scalars (bits, token counts) -> whole-job totals (e.g. bpp of a sharded batch)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.tolist()
