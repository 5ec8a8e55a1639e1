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


def reduce_metrics(metrics: dict[str, Tensor], group: dist.ProcessGroup | None = None) -> dict[str, Tensor]:
    """`sync_dist=True` mean-reduction of the logged scalars (core.py:243,265) batched into ONE small allreduce."""
    keys = sorted(metrics)
    flat = torch.stack([metrics[k].detach().float().reshape(()) for k in keys])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
        flat /= dist.get_world_size(group)
    return dict(zip(keys, flat.unbind()))
