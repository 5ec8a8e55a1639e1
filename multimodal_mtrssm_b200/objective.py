"""`likelihood` -- mirror of the reference's `models/objective.py:7-23` (same signature, same value).

CUDA tensors go through the fused streaming kernels of `librssm_rollout.so` (`rssm_gaussian_nll_fwd/_bwd`, one launch per
direction for ALL modalities handed to `likelihood_pairs`); there is no eager fallback for them.  CPU tensors (the
reference's function also works on them; host-side tests) use the closed form in torch.
"""

from __future__ import annotations

import ctypes as C
import math
from typing import List, Sequence

import torch
from torch import Tensor

from . import _lib

__all__ = ["likelihood", "likelihood_pairs"]

_workspaces: dict[tuple[int, int], Tensor] = {}


def _workspace(dev: torch.device) -> Tensor:
    """Zero-filled scratch of the forward kernel, one per (device, stream): calls leave it zero-filled (include/rssm_rollout.h)."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(dev).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(int(_lib.lib().rssm_gaussian_nll_workspace_bytes()), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws


def _check(predictions: Sequence[Tensor], targets: Sequence[Tensor]) -> int:
    if not 1 <= len(predictions) <= _lib.NLL_MAX_SEGMENTS or len(predictions) != len(targets):
        raise RuntimeError(f"gaussian_nll takes 1..{_lib.NLL_MAX_SEGMENTS} (prediction, target) pairs, got {len(predictions)}/{len(targets)}")
    dt = predictions[0].dtype
    if dt not in _lib.DTYPE_CODES:
        raise RuntimeError(f"gaussian_nll: prediction dtype {dt} is not one of fp32 / bf16 / fp16")
    for p, t in zip(predictions, targets):
        if not (p.is_cuda and t.is_cuda):
            raise RuntimeError("gaussian_nll kernels need CUDA tensors (no CPU fallback)")
        if p.dtype != dt or t.dtype != torch.float32 or p.shape != t.shape or not (p.is_contiguous() and t.is_contiguous()):
            raise RuntimeError(
                f"gaussian_nll: every prediction must be contiguous {dt}, every target contiguous fp32 of the same shape "
                f"(got {p.dtype}{tuple(p.shape)} / {t.dtype}{tuple(t.shape)})"
            )
    return _lib.DTYPE_CODES[dt]


def _pairs(predictions: Sequence[Tensor], targets: Sequence[Tensor], n_batch: Sequence[int], scale: float):  # noqa: ANN202
    arr = (_lib.NllPair * len(predictions))()
    for i, (p, t) in enumerate(zip(predictions, targets)):
        arr[i].prediction, arr[i].target = p.data_ptr(), t.data_ptr()
        arr[i].n_elems, arr[i].n_batch, arr[i].scale = p.numel(), int(n_batch[i]), float(scale)
    return arr


def _launch(name: str, *args) -> None:  # noqa: ANN002
    handle = _lib.lib()
    status = getattr(handle, name)(*args, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if status != 0:
        raise RuntimeError(f"{name} failed: {handle.rssm_last_error().decode()}")


@torch.library.custom_op("mtrssm_b200::gaussian_nll", mutates_args=())
def gaussian_nll_op(predictions: Sequence[Tensor], targets: Sequence[Tensor], n_batch: Sequence[int], scale: float) -> Tensor:
    """losses[i] = -mean log N(targets[i]; predictions[i], scale) summed over the event dims (objective.py:21-23)."""
    code = _check(predictions, targets)
    dev = predictions[0].device
    with torch.cuda.device(dev):
        losses = torch.empty(len(predictions), dtype=torch.float32, device=dev)
        arr = _pairs(predictions, targets, n_batch, scale)
        for i in range(len(predictions)):
            arr[i].loss = losses.data_ptr() + 4 * i
        ws = _workspace(dev)
        _launch("rssm_gaussian_nll_fwd", arr, len(predictions), code, C.c_void_p(ws.data_ptr()), ws.numel())
    return losses


@gaussian_nll_op.register_fake
def _(predictions, targets, n_batch, scale):  # noqa: ANN001, ANN202
    return predictions[0].new_empty(len(predictions), dtype=torch.float32)


@torch.library.custom_op("mtrssm_b200::gaussian_nll_bwd", mutates_args=())
def gaussian_nll_bwd_op(predictions: Sequence[Tensor], targets: Sequence[Tensor], n_batch: Sequence[int], scale: float,
                        d_losses: Tensor, target_grads: bool) -> List[Tensor]:
    """[d predictions..., (d targets... if target_grads)] of `gaussian_nll` for the upstream gradient d_losses [n_pairs]."""
    code = _check(predictions, targets)
    dev = predictions[0].device
    with torch.cuda.device(dev):
        d_losses = d_losses.float().contiguous()
        d_pred = [torch.empty_like(p) for p in predictions]
        d_tgt = [torch.empty_like(t) for t in targets] if target_grads else []
        arr = _pairs(predictions, targets, n_batch, scale)
        for i in range(len(predictions)):
            arr[i].d_loss = d_losses.data_ptr() + 4 * i
            arr[i].d_prediction = d_pred[i].data_ptr()
            arr[i].d_target = d_tgt[i].data_ptr() if target_grads else None
        _launch("rssm_gaussian_nll_bwd", arr, len(predictions), code)
    return d_pred + d_tgt


@gaussian_nll_bwd_op.register_fake
def _(predictions, targets, n_batch, scale, d_losses, target_grads):  # noqa: ANN001, ANN202
    return [torch.empty_like(p) for p in predictions] + ([torch.empty_like(t) for t in targets] if target_grads else [])


def _nll_setup(ctx, inputs, output) -> None:  # noqa: ANN001
    predictions, targets, n_batch, scale = inputs
    ctx.n, ctx.n_batch, ctx.scale = len(predictions), list(n_batch), scale
    ctx.target_grads = any(t.requires_grad for t in targets)
    ctx.save_for_backward(*predictions, *targets)


def _nll_backward(ctx, d_losses):  # noqa: ANN001, ANN202
    saved = ctx.saved_tensors
    grads = gaussian_nll_bwd_op(saved[: ctx.n], saved[ctx.n :], ctx.n_batch, ctx.scale, d_losses, ctx.target_grads)
    return list(grads[: ctx.n]), (list(grads[ctx.n :]) if ctx.target_grads else [None] * ctx.n), None, None


gaussian_nll_op.register_autograd(_nll_backward, setup_context=_nll_setup)


def likelihood_pairs(predictions: Sequence[Tensor], targets: Sequence[Tensor], event_ndims: int, scale: float = 1.0) -> Tensor:
    """`likelihood` of several (prediction, target) pairs -- the modalities of `compute_reconstruction_loss`
    (mrssm/mopoe_mrssm/core.py:294-303) -- in ONE kernel launch per direction.  Returns the losses as a [n_pairs] fp32 tensor."""
    preds, tgts, n_batch = [], [], []
    for p, t in zip(predictions, targets):
        if p.shape != t.shape:  # Normal(loc, scale).log_prob(value) broadcasts (objective.py:21-22)
            shape = torch.broadcast_shapes(p.shape, t.shape)
            p, t = p.expand(shape), t.expand(shape)
        preds.append(p.contiguous())
        tgts.append(t.float().contiguous())
        n_batch.append(max(1, math.prod(p.shape[: p.dim() - event_ndims])))
    return gaussian_nll_op(preds, tgts, n_batch, float(scale))


def likelihood(prediction: Tensor, target: Tensor, event_ndims: int, scale: float = 1.0) -> Tensor:
    """-mean over batch dims of log Independent(Normal(prediction, scale), event_ndims).log_prob(target).

    Closed form of objective.py:21-23: 0.5*||(x-mu)/scale||^2 + n*(log(scale) + 0.5*log(2*pi)) summed over the last
    `event_ndims` dims, then averaged -- one fused reduction instead of building distribution objects.
    """
    if prediction.is_cuda or target.is_cuda:
        return likelihood_pairs([prediction], [target], event_ndims, scale)[0]
    dims = tuple(range(-event_ndims, 0))
    n = math.prod(prediction.shape[-event_ndims:])
    sq = ((target - prediction) / scale).square().sum(dims)
    return (0.5 * sq + n * (math.log(scale) + 0.5 * math.log(2 * math.pi))).mean()
