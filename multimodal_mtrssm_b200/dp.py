"""Batch-sharded data parallelism for the rollout: one process per GPU, replicated parameters,
the batch split along dim 0, and ONE collective per step -- an allreduce of a single flat gradient bucket.

The rollout is independent per sequence, so there is no collective on the data path.  Parameters that never receive
a gradient (MoPoE-MMTRSSM's dummy `transition.*` and `l_posterior.*`, SURVEY.md §2.1 -- stock DDP errors on them)
are skipped: the bucket holds exactly the parameters whose `.grad` is not None, which is the same set on every
rank because every rank runs the same model code."""

from __future__ import annotations

from typing import Iterable, Sequence

import torch
import torch.distributed as dist
from torch import Tensor, nn


def bind_to_gpu_numa_node(device_index: int) -> bool:
    """Pins the calling process to the CPU cores NVML reports as local to GPU `device_index` (one process per GPU), so that
    pinned host buffers allocated afterwards are first-touched on that GPU's NUMA node and the per-step H2D copies of all ranks
    do not funnel through one socket's memory controllers / inter-socket link.  Returns False (and changes nothing) when NVML or
    the affinity call is unavailable; the data path never depends on it."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = device_index
        if visible:  # NVML enumerates physical devices
            entry = visible.split(",")[device_index].strip()
            if entry.isdigit():
                index = int(entry)
            else:
                pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByUUID(entry.encode()))
                return True
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
    except Exception:  # noqa: BLE001  (no NVML, no permission, restricted cpuset: keep the default placement)
        return False
    return True


def shard_batch(batch: Sequence[Tensor], rank: int, world: int) -> tuple[Tensor, ...]:
    """Equal contiguous shards along dim 0 (B must divide by world: loss means then average exactly)."""
    B = batch[0].shape[0]
    if B % world:
        msg = f"global batch {B} is not divisible by the world size {world}"
        raise ValueError(msg)
    per = B // world
    return tuple(t[rank * per:(rank + 1) * per] for t in batch)


class FlatGradBucket:
    """Averages gradients across ranks with one allreduce over one flat fp32 buffer."""

    def __init__(self, params: Iterable[nn.Parameter], group: dist.ProcessGroup | None = None) -> None:
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self._flat: Tensor | None = None

    def allreduce(self) -> int:
        """Average the existing `.grad`s in place; returns the number of elements reduced."""
        live = [p for p in self.params if p.grad is not None]
        if not live or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return 0
        n = sum(p.grad.numel() for p in live)
        if self._flat is None or self._flat.numel() != n or self._flat.device != live[0].grad.device:
            self._flat = torch.empty(n, dtype=torch.float32, device=live[0].grad.device)
        torch.cat([p.grad.reshape(-1).float() for p in live], out=self._flat)
        dist.all_reduce(self._flat, group=self.group)
        self._flat.div_(dist.get_world_size(self.group))
        off = 0
        for p in live:
            k = p.grad.numel()
            p.grad.copy_(self._flat[off:off + k].view_as(p.grad))
            off += k
        return n


def broadcast_parameters(module: nn.Module, src: int = 0, group: dist.ProcessGroup | None = None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized():
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def train_step(model: nn.Module, batch: Sequence[Tensor], optimizer: torch.optim.Optimizer, bucket: FlatGradBucket,
               clip: float | None = 10.0) -> dict[str, Tensor]:
    """One data-parallel step on this rank's shard: forward/backward, flat-bucket allreduce, clip, optimiser step.
    (`gradient_clip_val: 10`, AdamW: mopoe_*/configs/default.yaml:103-122.)"""
    optimizer.zero_grad(set_to_none=True)
    out = model.training_step(tuple(batch), 0)
    out["loss"].backward()
    bucket.allreduce()
    if clip is not None:
        torch.nn.utils.clip_grad_norm_([p for p in bucket.params if p.grad is not None], clip)
    optimizer.step()
    return out


class GraphedTrainStep:
    """`train_step` captured ONCE into a CUDA graph and replayed: the default.yaml models are launch-bound outside the fused
    rollout (a B = 8 step is a few hundred small encoder / decoder / optimiser launches), so replaying one graph removes the
    host from the step.  The batch is copied into static device buffers (the graph's inputs), the logged scalars come back as
    static tensors.  Everything inside is capture-safe: the rollout / likelihood kernels are plain launches on the capturing
    stream, the noise is `torch.rand` (graph-safe Philox offsets), the NCCL allreduce of the flat bucket is captured with the
    rest, gradient clipping stays on the device, and the optimiser must be built with `capturable=True`.

    Shapes are fixed at construction; a batch of another shape raises (build another instance for it)."""

    def __init__(self, model: nn.Module, example_batch: Sequence[Tensor], optimizer: torch.optim.Optimizer, bucket: FlatGradBucket,
                 clip: float | None = 10.0, autocast_dtype: torch.dtype | None = None, warmup: int = 3) -> None:
        if not all(t.is_cuda for t in example_batch):
            msg = "GraphedTrainStep needs CUDA tensors (CUDA graphs)"
            raise RuntimeError(msg)
        for group in optimizer.param_groups:
            if not group.get("capturable", False):
                msg = "GraphedTrainStep needs an optimizer built with capturable=True"
                raise RuntimeError(msg)
        self.model, self.optimizer, self.bucket, self.clip, self.autocast_dtype = model, optimizer, bucket, clip, autocast_dtype
        self.static_batch = tuple(t.clone() for t in example_batch)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up outside capture: lazy initialisations, cuDNN plans, optimiser state, .grad buffers
            for _ in range(max(1, warmup)):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = self._eager_step()

    def _eager_step(self) -> dict[str, Tensor]:
        self.optimizer.zero_grad(set_to_none=False)  # gradients keep their addresses: the graph accumulates into them
        with torch.autocast("cuda", dtype=self.autocast_dtype or torch.bfloat16, enabled=self.autocast_dtype is not None):
            out = self.model.training_step(self.static_batch, 0)
        out["loss"].backward()
        self.bucket.allreduce()
        if self.clip is not None:
            torch.nn.utils.clip_grad_norm_([p for p in self.bucket.params if p.grad is not None], self.clip)
        self.optimizer.step()
        return {k: v.detach() for k, v in out.items()}

    def __call__(self, batch: Sequence[Tensor]) -> dict[str, Tensor]:
        """Copies `batch` into the graph's input buffers and replays the step; returns the step's scalars (static tensors,
        overwritten by the next call)."""
        if len(batch) != len(self.static_batch):
            msg = f"batch has {len(batch)} tensors, the captured step takes {len(self.static_batch)}"
            raise ValueError(msg)
        for dst, src in zip(self.static_batch, batch):
            if dst.shape != src.shape:
                msg = f"batch shape {tuple(src.shape)} differs from the captured {tuple(dst.shape)}"
                raise ValueError(msg)
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


class PinnedPrefetcher:
    """Double-buffered host -> device input pipeline: step i+1's pinned host tensors are copied on a side stream while step i
    computes, so a copy-bound step costs max(copy, compute) instead of their sum.  `next()` returns the device tensors of the
    oldest pending batch (the consumer's stream waits for its copy) and `submit()` queues the following one; a buffer is
    reused only after the consumer's work on it was recorded as finished (`release`)."""

    def __init__(self, example: dict[str, Tensor], device: torch.device, depth: int = 2) -> None:
        self.device, self.depth = device, depth
        self.stream = torch.cuda.Stream(device)
        self.bufs = [{k: torch.empty(v.shape, dtype=v.dtype, device=device) for k, v in example.items()} for _ in range(depth)]
        self.copied = [torch.cuda.Event() for _ in range(depth)]
        self.free = [torch.cuda.Event() for _ in range(depth)]
        for e in self.free:
            e.record(torch.cuda.current_stream(device))
        self.head = self.tail = 0  # next slot to fill / next slot to hand out

    def submit(self, host: dict[str, Tensor]) -> None:
        if self.head - self.tail >= self.depth:
            msg = "PinnedPrefetcher: every buffer is in flight; call next()/release() first"
            raise RuntimeError(msg)
        slot = self.head % self.depth
        self.stream.wait_event(self.free[slot])
        with torch.cuda.stream(self.stream):
            for k, v in host.items():
                if not v.is_pinned():
                    msg = f"PinnedPrefetcher: host tensor {k!r} is not pinned (the copy would be synchronous)"
                    raise RuntimeError(msg)
                self.bufs[slot][k].copy_(v, non_blocking=True)
            self.copied[slot].record(self.stream)
        self.head += 1

    def next(self) -> tuple[int, dict[str, Tensor]]:
        if self.tail >= self.head:
            msg = "PinnedPrefetcher: nothing submitted"
            raise RuntimeError(msg)
        slot = self.tail % self.depth
        torch.cuda.current_stream(self.device).wait_event(self.copied[slot])
        self.tail += 1
        return slot, self.bufs[slot]

    def release(self, slot: int) -> None:
        """Marks the consumer's work queued so far as the last use of `slot`'s buffers."""
        self.free[slot].record(torch.cuda.current_stream(self.device))


def reduce_metrics(metrics: dict[str, Tensor], group: dist.ProcessGroup | None = None) -> dict[str, Tensor]:
    """`sync_dist=True` mean-reduction of the logged scalars (core.py:243,265) batched into ONE small allreduce."""
    keys = sorted(metrics)
    flat = torch.stack([metrics[k].detach().float().reshape(()) for k in keys])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
        flat /= dist.get_world_size(group)
    return dict(zip(keys, flat.unbind()))
