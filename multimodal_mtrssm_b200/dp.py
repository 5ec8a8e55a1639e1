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


class OverlappedGradBuckets:
    """Readiness-ordered gradient buckets: the allreduce of a bucket starts (asynchronously, on the collective's own stream) the
    moment the LAST gradient of that bucket has been accumulated, so it runs under the rest of the backward pass.  In the
    training step of the default models the gradients become ready as: decoders -> all rollout weights at once (one fused
    backward kernel) -> encoders; with the decoders' and the rollout's buckets already in flight only the encoders' bucket is
    exposed after `loss.backward()` returns.  (`FlatGradBucket` -- one allreduce after the backward -- stays for the CUDA-graph
    step, where the whole step is one replay.)

    The FIRST step runs synchronously and records, from per-parameter `post_accumulate_grad` hooks, WHICH parameters receive a
    gradient and in WHAT ORDER; parameters that never do (MoPoE-MMTRSSM's dummy `transition.*` / `l_posterior.*`) are left out,
    and the others are cut into buckets of about `bucket_bytes` along that order.  Every rank runs the same model code, so the
    order -- hence the bucket layout and the order of the collectives -- is the same on every rank; it is checked once by
    allreducing a checksum of the layout.  Use:

        buckets = OverlappedGradBuckets(model.parameters())
        ... loss.backward(); buckets.finish()        # every step; gradients are averaged in place when finish() returns
    """

    def __init__(self, params: Iterable[nn.Parameter], group: dist.ProcessGroup | None = None, bucket_bytes: int = 1 << 20) -> None:
        self.params = [p for p in params if p.requires_grad]
        self.group, self.bucket_bytes = group, bucket_bytes
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._order: list[int] = []          # first step: parameter indices in the order their gradients became ready
        self._bucket_of: dict[int, int] = {}  # parameter index -> bucket
        self._buckets: list[list[int]] = []
        self._flat: list[Tensor] = []
        self._pending: list[int] = []
        self._works: list[object | None] = []
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    @property
    def _active(self) -> bool:
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _on_grad(self, p: nn.Parameter) -> None:
        i = self._index[id(p)]
        if not self._buckets:  # recording step
            self._order.append(i)
            return
        b = self._bucket_of.get(i)
        if b is None:
            msg = "a parameter that had no gradient in the first step received one now: rebuild OverlappedGradBuckets"
            raise RuntimeError(msg)
        self._pending[b] -= 1
        if self._pending[b] == 0 and self._active:
            self._launch(b)

    def _launch(self, b: int) -> None:
        grads = [self.params[i].grad for i in self._buckets[b]]
        torch.cat([g.reshape(-1).float() for g in grads], out=self._flat[b])
        self._works[b] = dist.all_reduce(self._flat[b], group=self.group, async_op=True)

    def _build(self) -> None:
        order = list(dict.fromkeys(self._order))  # a shared parameter fires once per accumulation: keep the first
        cur, size = [], 0
        for i in order:
            cur.append(i)
            size += self.params[i].numel() * 4
            if size >= self.bucket_bytes:
                self._buckets.append(cur)
                cur, size = [], 0
        if cur:
            self._buckets.append(cur)
        for b, idx in enumerate(self._buckets):
            for i in idx:
                self._bucket_of[i] = b
        dev = self.params[order[0]].grad.device if order else torch.device("cpu")
        self._flat = [torch.empty(sum(self.params[i].numel() for i in idx), dtype=torch.float32, device=dev) for idx in self._buckets]
        self._works = [None] * len(self._buckets)
        if self._active:  # same layout on every rank?
            sig = torch.tensor([float(sum((k + 1) * (i + 1) for k, i in enumerate(order)) % 1000003), float(len(self._buckets))], device=dev)
            lo, hi = sig.clone(), sig.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
            if not torch.equal(lo, hi):
                msg = "ranks disagree on the gradient-readiness order; use FlatGradBucket"
                raise RuntimeError(msg)

    def finish(self) -> int:
        """Call after `backward()`: waits for the buckets in flight, averages, writes the gradients back.  Returns the number of
        elements reduced."""
        first = not self._buckets
        if first:
            self._build()
        if not self._active:
            self._pending = [len(idx) for idx in self._buckets]
            return 0
        world = dist.get_world_size(self.group)
        n = 0
        for b, idx in enumerate(self._buckets):
            if self._works[b] is None:  # recording step, or (defensively) a bucket whose last hook did not fire
                if not first and self._pending[b] != 0:
                    msg = f"bucket {b}: {self._pending[b]} gradients missing after backward"
                    raise RuntimeError(msg)
                self._launch(b)
            self._works[b].wait()
            self._works[b] = None
            flat = self._flat[b].div_(world)
            off = 0
            for i in idx:
                g = self.params[i].grad
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
            n += off
        self._pending = [len(idx) for idx in self._buckets]
        return n

    allreduce = finish  # drop-in for FlatGradBucket in `train_step`

    def remove_hooks(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


class _DeviceView:
    """`__cuda_array_interface__` over raw device memory, so torch can wrap a region of the peer-mapped allocation."""

    def __init__(self, ptr: int, n_floats: int) -> None:
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


class P2pGradAllreduce:
    """One-shot mean-allreduce of a small fp32 gradient bucket over NVLink / NVSwitch peer memory
    (`rssm_p2p_allreduce_mean`, csrc/p2p_allreduce.cu): every rank keeps the bucket in a cudaIpc-mapped allocation, a single
    kernel exchanges one flag per peer and reads all ranks' buckets straight from peer memory.  For the rollout's 66 KB bucket
    this replaces a latency-bound NCCL allreduce (~37 us per step at 8 GPUs) on the compute stream.

        comm = P2pGradAllreduce(n_floats)            # collective: every rank of `group` (one box, <= 8 ranks) must call it
        ... backward writes the weight gradients into `flat` (ordinary device memory) ...
        comm.allreduce(step, flat)                   # flat = mean over ranks, in place; `step` = 0, 1, 2, ... the same on every rank

    The handles travel through `torch.distributed.all_gather_object` once, at construction.  Raises if peer access between the
    ranks' GPUs cannot be had (the caller then keeps NCCL)."""

    def __init__(self, n_floats: int, group: dist.ProcessGroup | None = None, device: torch.device | None = None,
                 timeout_ms: int = 5000) -> None:
        import ctypes as C

        from . import _lib

        if not dist.is_initialized():
            raise RuntimeError("P2pGradAllreduce needs an initialised torch.distributed process group")
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.P2P_MAX_RANKS:
            raise RuntimeError(f"P2pGradAllreduce: at most {_lib.P2P_MAX_RANKS} ranks of one box (got {self.world})")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.n, self.timeout_ms, self._lib, self._C = n_floats, timeout_ms, _lib, C
        lib = _lib.lib()
        with torch.cuda.device(self.device):
            nbytes = lib.rssm_p2p_region_bytes(n_floats)
            own = C.c_void_p()
            if lib.rssm_p2p_alloc(nbytes, C.byref(own)):
                raise RuntimeError(f"rssm_p2p_alloc failed: {lib.rssm_last_error().decode()}")
            self._own = own.value
            handle = C.create_string_buffer(64)
            if lib.rssm_p2p_export(self._own, handle):
                raise RuntimeError(f"rssm_p2p_export failed: {lib.rssm_last_error().decode()}")
            handles: list = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self.comm = _lib.P2pComm(world=self.world, rank=self.rank, n=n_floats)
            self._imported = []
            errors = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.comm.regions[r] = self._own
                    continue
                peer = C.c_void_p()
                if lib.rssm_p2p_import(h, C.byref(peer)):
                    errors.append(f"rank {r}: {lib.rssm_last_error().decode()}")
                    continue
                self.comm.regions[r] = peer.value
                self._imported.append(peer.value)
            ok = torch.tensor([0.0 if errors else 1.0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # every rank keeps or drops the peer path together
            if float(ok) == 0.0:
                self.close()
                raise RuntimeError("P2pGradAllreduce: peer mapping failed" + (": " + "; ".join(errors) if errors else " on another rank"))
            self._buckets = [torch.as_tensor(_DeviceView(lib.rssm_p2p_bucket(self._own, n_floats, k), n_floats), device=self.device)
                             for k in range(2)]

    def bucket(self, step: int) -> Tensor:
        return self._buckets[step & 1]

    def allreduce(self, step: int, src: Tensor | None, out: Tensor | None = None) -> None:
        """out[0:n] = mean over ranks, on the current stream of this rank's device.  `src`: this rank's gradients (ordinary device
        memory; copied into the peer-mapped bucket of this step first) -- or None when the caller filled `bucket(step)` in place.
        `out` defaults to `src` (in place)."""
        out = src if out is None else out
        for t in (src, out):
            if t is not None and (t.numel() != self.n or t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device):
                raise RuntimeError("P2pGradAllreduce.allreduce: contiguous fp32 tensors of the bucket's size on this device, please")
        if out is None:
            raise RuntimeError("P2pGradAllreduce.allreduce: give `src` or `out`")
        with torch.cuda.device(self.device):
            C, lib = self._C, self._lib.lib()
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            if lib.rssm_p2p_allreduce_mean(C.byref(self.comm), step, C.c_void_p(src.data_ptr()) if src is not None else None,
                                           C.c_void_p(out.data_ptr()), self.timeout_ms, stream):
                raise RuntimeError(f"rssm_p2p_allreduce_mean failed: {lib.rssm_last_error().decode()}")

    def check(self) -> None:
        """Synchronises and raises if a call gave up waiting for a peer."""
        if self._lib.lib().rssm_p2p_status(self._C.byref(self.comm)) != 0:
            raise RuntimeError("P2pGradAllreduce: a peer did not arrive within the timeout")

    def close(self) -> None:
        lib = self._lib.lib()
        for p in getattr(self, "_imported", []):
            lib.rssm_p2p_close(p)
        self._imported = []
        self._buckets = []
        if getattr(self, "_own", None):
            lib.rssm_p2p_free(self._own)
            self._own = None


def broadcast_parameters(module: nn.Module, src: int = 0, group: dist.ProcessGroup | None = None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized():
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def train_step(model: nn.Module, batch: Sequence[Tensor], optimizer: torch.optim.Optimizer, bucket: "FlatGradBucket | OverlappedGradBuckets",
               clip: float | None = 10.0) -> dict[str, Tensor]:
    """One data-parallel step on this rank's shard: forward/backward, gradient allreduce, clip, optimiser step.
    (`gradient_clip_val: 10`, AdamW: mopoe_*/configs/default.yaml:103-122.)  With `OverlappedGradBuckets` the allreduce of the
    decoders' and the rollout's gradients runs under the encoders' backward; `FlatGradBucket` reduces once after it."""
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
