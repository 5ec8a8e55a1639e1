"""torch custom ops (with autograd) over the C-ABI rollout library.

Ops (namespace `mtrssm_b200`):
  mrssm_rollout / mrssm_rollout_bwd / mrssm_imagine      -- MoPoE-MRSSM
  mtrssm_rollout / mtrssm_rollout_bwd / mtrssm_imagine   -- MoPoE-MMTRSSM

They are thin: allocate outputs with torch, pass raw device pointers to `librssm_rollout.so` on the current
CUDA stream.  CUDA only -- a CPU tensor raises; there is no eager fallback.

Weight order of the `weights` list follows `_lib.MR_WEIGHT_FIELDS` / `_lib.MT_WEIGHT_FIELDS`.
"""

from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
from torch import Tensor

from . import _lib
from ._lib import ptr

__all__ = ["mrssm_rollout", "mrssm_imagine", "mtrssm_rollout", "mtrssm_imagine", "kl_path_weights", "obs_projection"]

KL_BALANCE_ALPHA = 0.8


def kl_path_weights(use_balancing: bool) -> tuple[float, float]:
    """(weight of d KL/d posterior, weight of d KL/d prior): balanced = (1-a, a), plain = (1, 1)."""
    return (1.0 - KL_BALANCE_ALPHA, KL_BALANCE_ALPHA) if use_balancing else (1.0, 1.0)


def _c(t: Optional[Tensor]) -> Optional[Tensor]:
    """contiguous fp32 (inputs may arrive in half precision under autocast; state and I/O of the kernels are fp32)"""
    return None if t is None else t.float().contiguous()


def _on_device(*tensors):  # noqa: ANN002, ANN202
    """Device guard of one op call: every tensor must live on ONE CUDA device (the first tensor's); the body then runs with that
    device current, so `torch.empty(device=...)`, `torch.cuda.current_stream()` and the kernel launch all agree even when the
    caller's current device is another GPU.  CPU tensors raise: there is no eager fallback."""
    flat = []
    for t in tensors:
        if isinstance(t, Tensor):
            flat.append(t)
        elif isinstance(t, (list, tuple)):
            flat.extend(x for x in t if isinstance(x, Tensor))
    dev = flat[0].device
    if dev.type != "cuda":
        raise RuntimeError(f"the fused rollout ops are CUDA-only (there is no CPU fallback); got a tensor on {dev}")
    for t in flat:
        if t.device != dev:
            raise RuntimeError(f"all tensors of a fused rollout call must be on one device: got {dev} and {t.device}")
    return torch.cuda.device(dev)


def _fill(struct, names: Sequence[str], tensors: Sequence[Optional[Tensor]]):  # noqa: ANN001
    for n, t in zip(names, tensors):
        setattr(struct, n, ptr(t))
    return struct


# =================================================================================================
# MoPoE-MRSSM
# =================================================================================================
def _mr_dims(actions: Tensor, K: int, precision: int, D: int = 32, unimodal: bool = False) -> _lib.MrssmDims:
    B, T, A = actions.shape
    return _lib.MrssmDims(B=B, T=T, A=A, E=64, D=D, H=D, C=16 // K, K=K, precision=precision, unimodal=int(unimodal))


def _mr_wide(D: int) -> bool:
    """deterministic_size of the wide family (persistent tcgen05 kernels; include/rssm_rollout.h)"""
    return D != 32


def _mr_check(weights: Sequence[Tensor], embed_a: Tensor, h0: Tensor, z0: Tensor) -> int:
    """Validates the size family; returns deterministic_size."""
    if len(weights) != len(_lib.MR_WEIGHT_FIELDS):
        raise RuntimeError(f"expected {len(_lib.MR_WEIGHT_FIELDS)} weight tensors, got {len(weights)}")
    D = h0.shape[-1]
    ok_d = D == 32 or (D % 128 == 0 and 128 <= D <= 512)
    if embed_a.shape[-1] != 64 or not ok_d or z0.shape[-1] != 16 or tuple(weights[4].shape) != (3 * D, D):
        raise RuntimeError(
            "fused MRSSM rollout supports deterministic_size = hidden_size = 32 or a multiple of 128 up to 512, "
            f"obs_embed_size = 64, class_size*category_size = 16 (got embed {embed_a.shape[-1]}, deter {h0.shape[-1]}, "
            f"stoch {z0.shape[-1]}, weight_ih {tuple(weights[4].shape)})"
        )
    return D


_WIDE_PLANES = 10  # record planes per step of the wide family (csrc/wide_common.cuh)


def _mr_saved_shape(B: int, T: int, D: int) -> tuple[int, ...]:
    """Shape of the opaque saved record.  Wide family: bf16 elements covering the planes [T][plane][block][D/8][128][8] followed
    by the fp32 audio / vision logits [blocks*128][T][32] (rssm_mrssm_saved_bytes)."""
    if not _mr_wide(D):
        return (B, T, _lib.MRSSM_SAVED_FLOATS)
    blocks = (B + 127) // 128
    return (T * _WIDE_PLANES * blocks * D * 128 + blocks * 128 * T * 32 * 2,)


_DEBUG_LAST_WORKSPACE: dict[bool, Tensor] = {}  # RSSM_WIDE_TIMING only: lets a profiling script read the phase timestamps


def _mr_workspace(dims: _lib.MrssmDims, backward: bool, dev: torch.device) -> Optional[Tensor]:
    n = _lib.mrssm_workspace_bytes(dims, backward)
    ws = torch.empty((n + 1) // 2, device=dev, dtype=torch.bfloat16) if n else None
    if ws is not None and os.environ.get("RSSM_WIDE_TIMING"):
        _DEBUG_LAST_WORKSPACE[backward] = ws
    return ws


@torch.library.custom_op("mtrssm_b200::mrssm_rollout", mutates_args=())
def mrssm_rollout_op(
    weights: Sequence[Tensor], actions: Tensor, embed_a: Tensor, embed_v: Tensor, h0: Tensor, z0: Tensor,
    u_post: Tensor, u_prior: Optional[Tensor], K: int, precision: int, kl_wq: float, kl_wp: float, save: bool,
    unimodal: bool,
) -> List[Tensor]:
    with _on_device(actions, weights, embed_a, embed_v, h0, z0, u_post, u_prior):
        D = _mr_check(weights, embed_a, h0, z0)
        B, T, _ = actions.shape
        dev = actions.device
        dims = _mr_dims(actions, K, precision, D, unimodal)
        feature = torch.empty(B, T, D + 16, device=dev)
        prior_probs = torch.empty(B, T, 16 // K, K, device=dev)
        post_probs = torch.empty_like(prior_probs)
        prior_stoch = torch.empty(B, T, 16, device=dev) if u_prior is not None else torch.empty(0, device=dev)
        kl = torch.empty(B, T, device=dev)
        saved = torch.empty(_mr_saved_shape(B, T, D), device=dev, dtype=_lib.record_dtype(precision)) if save else torch.empty(0, device=dev)
        lib_bytes = _lib.mrssm_saved_bytes(dims)  # 0: unsupported sizes / precision -- the launch below reports why
        if save and lib_bytes and saved.numel() * saved.element_size() != lib_bytes:
            raise RuntimeError(f"saved record size mismatch with the library: {saved.numel() * saved.element_size()} vs {lib_bytes} bytes")
        workspace = _mr_workspace(dims, False, dev)
        w = _fill(_lib.MrssmWeights(), _lib.MR_WEIGHT_FIELDS, weights)
        inp = _fill(_lib.MrssmInputs(), "actions embed_a embed_v h0 z0 u_post u_prior".split(),
                    (actions, embed_a, embed_v, h0, z0, u_post, u_prior))
        out = _fill(_lib.MrssmOutputs(), "feature prior_probs post_probs prior_stoch kl saved workspace".split(),
                    (feature, prior_probs, post_probs, prior_stoch if u_prior is not None else None, kl, saved if save else None,
                     workspace))
        out.workspace_bytes = 0 if workspace is None else workspace.numel() * 2
        _lib.call("rssm_mrssm_rollout_fwd", dims, w, inp, out)
        return [feature, prior_probs, post_probs, prior_stoch, kl, saved]


@mrssm_rollout_op.register_fake
def _(weights, actions, embed_a, embed_v, h0, z0, u_post, u_prior, K, precision, kl_wq, kl_wp, save, unimodal):  # noqa: ANN001
    B, T, _ = actions.shape
    D = h0.shape[-1]
    e = actions.new_empty
    return [e(B, T, D + 16), e(B, T, 16 // K, K), e(B, T, 16 // K, K), e(B, T, 16) if u_prior is not None else e(0), e(B, T),
            e(_mr_saved_shape(B, T, D), dtype=_lib.record_dtype(precision)) if save else e(0)]


@torch.library.custom_op("mtrssm_b200::mrssm_rollout_bwd", mutates_args=())
def mrssm_rollout_bwd_op(
    weights: Sequence[Tensor], actions: Tensor, embed_a: Tensor, embed_v: Tensor, h0: Tensor, z0: Tensor,
    feature: Tensor, prior_probs: Tensor, post_probs: Tensor, saved: Tensor,
    d_feature: Optional[Tensor], d_prior_probs: Optional[Tensor], d_post_probs: Optional[Tensor],
    d_prior_stoch: Optional[Tensor], d_kl: Optional[Tensor], K: int, precision: int, kl_wq: float, kl_wp: float,
    unimodal: bool,
) -> List[Tensor]:
    with _on_device(actions, weights, embed_a, embed_v, h0, z0, feature, prior_probs, post_probs, saved, d_feature, d_prior_probs, d_post_probs, d_prior_stoch, d_kl):
        B, T, A = actions.shape
        dev = actions.device
        D = h0.shape[-1]
        dims = _mr_dims(actions, K, precision, D, unimodal)
        if d_feature is None:
            d_feature = torch.zeros_like(feature)
        sizes = [t.numel() for t in weights]
        flat = torch.zeros(sum(sizes), device=dev)
        gws = [g.view_as(t) for g, t in zip(flat.split(sizes), weights)]
        d_actions = torch.empty(B, T, A, device=dev)
        d_embed_a = torch.empty(B, T, 64, device=dev)
        d_embed_v = torch.empty(B, T, 64, device=dev)
        d_h0 = torch.empty(B, D, device=dev)
        d_z0 = torch.empty(B, 16, device=dev)
        # default family: pre-activation gradient record; wide family: the library keeps its gradient planes in the workspace
        dpre = None if _mr_wide(D) else torch.empty(B, T, _lib.MRSSM_DPRE_FLOATS, device=dev, dtype=_lib.record_dtype(precision))
        workspace = _mr_workspace(dims, True, dev)
        w = _fill(_lib.MrssmWeights(), _lib.MR_WEIGHT_FIELDS, weights)
        gw = _fill(_lib.MrssmWeightGrads(), _lib.MR_WEIGHT_FIELDS, gws)
        inp = _fill(_lib.MrssmInputs(), "actions embed_a embed_v h0 z0".split(), (actions, embed_a, embed_v, h0, z0))
        out = _fill(_lib.MrssmOutputs(), "feature prior_probs post_probs saved".split(), (feature, prior_probs, post_probs, saved))
        up = _fill(_lib.MrssmUpstream(), "d_feature d_prior_probs d_post_probs d_prior_stoch d_kl".split(),
                   (_c(d_feature), _c(d_prior_probs), _c(d_post_probs), _c(d_prior_stoch), _c(d_kl)))
        up.kl_wq, up.kl_wp = kl_wq, kl_wp
        gin = _fill(_lib.MrssmInputGrads(), "d_actions d_embed_a d_embed_v d_h0 d_z0 dpre workspace".split(),
                    (d_actions, d_embed_a, d_embed_v, d_h0, d_z0, dpre, workspace))
        gin.workspace_bytes = 0 if workspace is None else workspace.numel() * 2
        _lib.call("rssm_mrssm_rollout_bwd", dims, w, inp, out, up, gin, gw)
        return [flat, d_actions, d_embed_a, d_embed_v, d_h0, d_z0]  # flat = all weight grads, split by the caller


@mrssm_rollout_bwd_op.register_fake
def _(weights, actions, embed_a, embed_v, h0, z0, feature, prior_probs, post_probs, saved, d_feature, d_prior_probs,  # noqa: ANN001
      d_post_probs, d_prior_stoch, d_kl, K, precision, kl_wq, kl_wp, unimodal):
    return [actions.new_empty(sum(t.numel() for t in weights)), torch.empty_like(actions), torch.empty_like(embed_a),
            torch.empty_like(embed_v), torch.empty_like(h0), torch.empty_like(z0)]


def _mr_setup(ctx, inputs, output) -> None:  # noqa: ANN001
    weights, actions, embed_a, embed_v, h0, z0, u_post, u_prior, K, precision, kl_wq, kl_wp, save, unimodal = inputs
    feature, prior_probs, post_probs, prior_stoch, kl, saved = output
    if not save:
        raise RuntimeError("mrssm_rollout was called with save=False but a gradient is required")
    ctx.set_materialize_grads(False)
    ctx.nw = len(weights)
    ctx.has_prior_stoch = u_prior is not None
    ctx.cfg = (K, precision, kl_wq, kl_wp, unimodal)
    ctx.save_for_backward(*weights, actions, embed_a, embed_v, h0, z0, feature, prior_probs, post_probs, saved)


def _mr_backward(ctx, grads):  # noqa: ANN001
    d_feature, d_prior_probs, d_post_probs, d_prior_stoch, d_kl, _ = grads
    saved_t = ctx.saved_tensors
    weights, rest = list(saved_t[: ctx.nw]), saved_t[ctx.nw:]
    actions, embed_a, embed_v, h0, z0, feature, prior_probs, post_probs, saved = rest
    K, precision, kl_wq, kl_wp, unimodal = ctx.cfg
    res = mrssm_rollout_bwd_op(
        weights, actions, embed_a, embed_v, h0, z0, feature, prior_probs, post_probs, saved, d_feature, d_prior_probs,
        d_post_probs, d_prior_stoch if ctx.has_prior_stoch else None, d_kl, K, precision, kl_wq, kl_wp, unimodal,
    )
    flat, d_actions, d_ea, d_ev, d_h0, d_z0 = res
    gws = [g.view_as(w) for g, w in zip(flat.split([w.numel() for w in weights]), weights)]
    return gws, d_actions, d_ea, d_ev, d_h0, d_z0, None, None, None, None, None, None, None, None


mrssm_rollout_op.register_autograd(_mr_backward, setup_context=_mr_setup)


def mrssm_rollout(
    weights: Sequence[Tensor], *, actions: Tensor, embed_a: Tensor, embed_v: Tensor, h0: Tensor, z0: Tensor,
    u_post: Tensor, u_prior: Optional[Tensor] = None, class_size: int = 4, precision: int = _lib.PRECISION_FP32,
    use_kl_balancing: bool = True, unimodal: bool = False,
) -> dict[str, Tensor]:
    """Fused MoPoE-MRSSM rollout_representation on encoder outputs (mrssm/mopoe_mrssm/core.py:184-260).

    `unimodal=True` is BaseRSSM.rollout_representation (core.py:137-168): the posterior is the FIRST head (`au_*` weights on
    `embed_a`, i.e. `representation.rnn_to_post_projector`) with no fusion; `embed_v` and the `vi_*` weights are not used by the
    result (pass the audio ones) and get zero gradients.

    Returns feature [B,T,48] = [deter | post_stoch], prior_probs / post_probs [B,T,C,K], prior_stoch
    [B,T,16] (only when `u_prior` is given) and kl [B,T] = sum_c KL(post_c || prior_c) whose gradient is
    routed with the KL-balancing weights chosen here.
    """
    save = torch.is_grad_enabled() and any(t.requires_grad for t in (*weights, actions, embed_a, embed_v, h0, z0))
    wq, wp = kl_path_weights(use_kl_balancing)
    feature, prior_probs, post_probs, prior_stoch, kl, _ = mrssm_rollout_op(
        [_c(w) for w in weights], _c(actions), _c(embed_a), _c(embed_v), _c(h0), _c(z0), _c(u_post), _c(u_prior),
        class_size, precision, wq, wp, save, unimodal,
    )
    return {"feature": feature, "prior_probs": prior_probs, "post_probs": post_probs,
            "prior_stoch": prior_stoch if u_prior is not None else None, "kl": kl}


@torch.library.custom_op("mtrssm_b200::mrssm_imagine", mutates_args=())
def mrssm_imagine_op(weights: Sequence[Tensor], actions: Tensor, h0: Tensor, z0: Tensor, u: Tensor, K: int,
                     precision: int) -> List[Tensor]:
    with _on_device(actions, weights, h0, z0, u):
        B, T, _ = actions.shape
        dev = actions.device
        D = h0.shape[-1]
        dims = _mr_dims(actions, K, precision, D)
        feature = torch.empty(B, T, D + 16, device=dev)
        probs = torch.empty(B, T, 16 // K, K, device=dev)
        workspace = _mr_workspace(dims, False, dev)
        w = _fill(_lib.MrssmWeights(), _lib.MR_WEIGHT_FIELDS, weights)
        inp = _fill(_lib.MrssmInputs(), "actions h0 z0 u_prior".split(), (actions, h0, z0, u))
        out = _fill(_lib.MrssmOutputs(), "feature prior_probs workspace".split(), (feature, probs, workspace))
        out.workspace_bytes = 0 if workspace is None else workspace.numel() * 2
        _lib.call("rssm_mrssm_imagine_fwd", dims, w, inp, out)
        return [feature, probs]


@mrssm_imagine_op.register_fake
def _(weights, actions, h0, z0, u, K, precision):  # noqa: ANN001
    B, T, _ = actions.shape
    return [actions.new_empty(B, T, h0.shape[-1] + 16), actions.new_empty(B, T, 16 // K, K)]


def mrssm_imagine(weights: Sequence[Tensor], *, actions: Tensor, h0: Tensor, z0: Tensor, u: Tensor, class_size: int = 4,
                  precision: int = _lib.PRECISION_FP32) -> dict[str, Tensor]:
    """Fused BaseRSSM.rollout_transition (core.py:170-185).  Forward only (the reference calls it under no_grad)."""
    with torch.no_grad():
        feature, probs = mrssm_imagine_op([_c(w) for w in weights], _c(actions), _c(h0), _c(z0), _c(u), class_size, precision)
    return {"feature": feature, "probs": probs}


# =================================================================================================
# MoPoE-MMTRSSM
# =================================================================================================
def _mt_dims(actions: Tensor, KL: int, KH: int, l_tau: float, h_tau: float, precision: int, obs_projected: bool = False) -> _lib.MtrssmDims:
    B, T, A = actions.shape
    return _lib.MtrssmDims(B=B, T=T, A=A, E=64, HD=32, LD=32, HH=32, HR=32, CL=16 // KL, KL=KL, CH=16 // KH, KH=KH,
                           l_tau=l_tau, h_tau=h_tau, precision=precision, obs_projected=int(obs_projected))


_MT_STATE = "deter_h0 deter_l0 hidden_h0 hidden_l0 stoch_h0 stoch_l0".split()


def _mt_check(weights: Sequence[Tensor], embed_a: Tensor, state: Sequence[Tensor], obs_projected: bool = False) -> None:
    if len(weights) != len(_lib.MT_WEIGHT_FIELDS):
        raise RuntimeError(f"expected {len(_lib.MT_WEIGHT_FIELDS)} weight tensors, got {len(weights)}")
    shapes = [s.shape[-1] for s in state]
    if obs_projected and embed_a.shape[-1] != 32:
        raise RuntimeError(f"obs_projected=True takes the pre-multiplied partials e @ W1[:, 32:].T of width 32, got {embed_a.shape[-1]}")
    if (not obs_projected and embed_a.shape[-1] != 64) or shapes != [32, 32, 32, 32, 16, 16] or tuple(weights[8].shape) != (32, 32):
        raise RuntimeError(
            "fused MMTRSSM rollout supports hd_dim = ld_dim = 32, hs_dim = ls_dim = 16, head hidden 32, obs_embed_size 64 "
            f"(got embed {embed_a.shape[-1]}, state widths {shapes}, l_prior.0 {tuple(weights[8].shape)})"
        )


@torch.library.custom_op("mtrssm_b200::mtrssm_rollout", mutates_args=())
def mtrssm_rollout_op(
    weights: Sequence[Tensor], actions: Tensor, embed_a: Tensor, embed_v: Tensor, state: Sequence[Tensor],
    u_post_l: Tensor, u_post_h: Tensor, u_prior_l: Optional[Tensor], u_prior_h: Optional[Tensor],
    KL: int, KH: int, l_tau: float, h_tau: float, precision: int, kl_wq: float, kl_wp: float, save: bool,
    obs_projected: bool,  # no default: the dispatcher strips trailing default-valued arguments, which changes the backward's arity
) -> List[Tensor]:
    with _on_device(actions, weights, embed_a, embed_v, state, u_post_l, u_post_h, u_prior_l, u_prior_h):
        _mt_check(weights, embed_a, state, obs_projected)
        B, T, _ = actions.shape
        dev = actions.device
        e = lambda *s: torch.empty(*s, device=dev)  # noqa: E731
        feature, hidden_h, hidden_l = e(B, T, 96), e(B, T, 32), e(B, T, 32)
        prior_h, post_h = e(B, T, 16 // KH, KH), e(B, T, 16 // KH, KH)
        prior_l, post_l = e(B, T, 16 // KL, KL), e(B, T, 16 // KL, KL)
        has_prior = u_prior_l is not None
        pz_h, pz_l = (e(B, T, 16), e(B, T, 16)) if has_prior else (e(0), e(0))
        kl_l, kl_h = e(B, T), e(B, T)
        saved = (torch.empty(_lib.mtrssm_saved_rows(B, precision), T, _lib.mtrssm_saved_elems(precision), device=dev,
                             dtype=_lib.record_dtype(precision)) if save else e(0))
        w = _fill(_lib.MtrssmWeights(), _lib.MT_WEIGHT_FIELDS, weights)
        inp = _fill(_lib.MtrssmInputs(), ["actions", "embed_a", "embed_v", *_MT_STATE, "u_post_l", "u_post_h", "u_prior_l", "u_prior_h"],
                    (actions, embed_a, embed_v, *state, u_post_l, u_post_h, u_prior_l, u_prior_h))
        out = _fill(
            _lib.MtrssmOutputs(),
            "feature hidden_h hidden_l prior_probs_h prior_probs_l post_probs_h post_probs_l prior_stoch_h prior_stoch_l kl_l kl_h saved".split(),
            (feature, hidden_h, hidden_l, prior_h, prior_l, post_h, post_l, pz_h if has_prior else None, pz_l if has_prior else None,
             kl_l, kl_h, saved if save else None),
        )
        _lib.call("rssm_mtrssm_rollout_fwd", _mt_dims(actions, KL, KH, l_tau, h_tau, precision, obs_projected), w, inp, out)
        return [feature, hidden_h, hidden_l, prior_h, prior_l, post_h, post_l, pz_h, pz_l, kl_l, kl_h, saved]


@mtrssm_rollout_op.register_fake
def _(weights, actions, embed_a, embed_v, state, u_post_l, u_post_h, u_prior_l, u_prior_h, KL, KH, l_tau, h_tau, precision,  # noqa: ANN001
      kl_wq, kl_wp, save, obs_projected):
    B, T, _ = actions.shape
    e = actions.new_empty
    pz = e(B, T, 16) if u_prior_l is not None else e(0)
    return [e(B, T, 96), e(B, T, 32), e(B, T, 32), e(B, T, 16 // KH, KH), e(B, T, 16 // KL, KL), e(B, T, 16 // KH, KH),
            e(B, T, 16 // KL, KL), pz, torch.empty_like(pz), e(B, T), e(B, T),
            e(_lib.mtrssm_saved_rows(B, precision), T, _lib.mtrssm_saved_elems(precision), dtype=_lib.record_dtype(precision)) if save else e(0)]


def _hidden_pair(d_hh: Optional[Tensor], d_hl: Optional[Tensor], B: int, T: int, dev: torch.device):  # noqa: ANN202
    """Upstream gradients of the hidden_h / hidden_l outputs: the kernels take both or neither."""
    if d_hh is None and d_hl is None:
        return None, None
    z = lambda t: torch.zeros(B, T, 32, device=dev) if t is None else t  # noqa: E731
    return z(d_hh), z(d_hl)


@torch.library.custom_op("mtrssm_b200::mtrssm_rollout_bwd", mutates_args=())
def mtrssm_rollout_bwd_op(
    weights: Sequence[Tensor], actions: Tensor, embed_a: Tensor, embed_v: Tensor, state: Sequence[Tensor],
    feature: Tensor, prior_h: Tensor, prior_l: Tensor, post_h: Tensor, post_l: Tensor, saved: Tensor,
    d_feature: Optional[Tensor], d_prior_h: Optional[Tensor], d_prior_l: Optional[Tensor], d_post_h: Optional[Tensor],
    d_post_l: Optional[Tensor], d_pz_h: Optional[Tensor], d_pz_l: Optional[Tensor], d_kl_l: Optional[Tensor],
    d_kl_h: Optional[Tensor], d_hidden_h: Optional[Tensor], d_hidden_l: Optional[Tensor],
    KL: int, KH: int, l_tau: float, h_tau: float, precision: int, kl_wq: float, kl_wp: float,
    obs_projected: bool,  # no default: the dispatcher strips trailing default-valued arguments, which changes the backward's arity
) -> List[Tensor]:
    with _on_device(actions, weights, embed_a, embed_v, state, feature, prior_h, prior_l, post_h, post_l, saved, d_feature, d_prior_h, d_prior_l, d_post_h, d_post_l, d_pz_h, d_pz_l, d_kl_l, d_kl_h,
                    d_hidden_h, d_hidden_l):
        B, T, A = actions.shape
        dev = actions.device
        if d_feature is None:
            d_feature = torch.zeros_like(feature)
        d_hidden_h, d_hidden_l = _hidden_pair(d_hidden_h, d_hidden_l, B, T, dev)
        sizes = [t.numel() for t in weights]
        flat = torch.zeros(sum(sizes), device=dev)
        gws = [g.view_as(t) for g, t in zip(flat.split(sizes), weights)]
        e = lambda *s: torch.empty(*s, device=dev)  # noqa: E731
        EW = 32 if obs_projected else 64  # width of the embedding inputs and of their gradients
        d_actions, d_ea, d_ev = e(B, T, A), e(B, T, EW), e(B, T, EW)
        d_state = [e(B, 32), e(B, 32), e(B, 32), e(B, 32), e(B, 16), e(B, 16)]
        dpre = torch.empty(B, T, _lib.MTRSSM_DPRE_FLOATS, device=dev, dtype=_lib.record_dtype(precision))
        w = _fill(_lib.MtrssmWeights(), _lib.MT_WEIGHT_FIELDS, weights)
        gw = _fill(_lib.MtrssmWeightGrads(), _lib.MT_WEIGHT_FIELDS, gws)
        inp = _fill(_lib.MtrssmInputs(), ["actions", "embed_a", "embed_v", *_MT_STATE], (actions, embed_a, embed_v, *state))
        out = _fill(_lib.MtrssmOutputs(), "feature prior_probs_h prior_probs_l post_probs_h post_probs_l saved".split(),
                    (feature, prior_h, prior_l, post_h, post_l, saved))
        up = _fill(
            _lib.MtrssmUpstream(),
            ("d_feature d_prior_probs_h d_prior_probs_l d_post_probs_h d_post_probs_l d_prior_stoch_h d_prior_stoch_l d_kl_l d_kl_h "
             "d_hidden_h d_hidden_l").split(),
            tuple(_c(t) for t in (d_feature, d_prior_h, d_prior_l, d_post_h, d_post_l, d_pz_h, d_pz_l, d_kl_l, d_kl_h, d_hidden_h,
                                  d_hidden_l)),
        )
        up.kl_wq, up.kl_wp = kl_wq, kl_wp
        gin = _fill(_lib.MtrssmInputGrads(),
                    ["d_actions", "d_embed_a", "d_embed_v", *("d_" + n for n in _MT_STATE), "dpre"],
                    (d_actions, d_ea, d_ev, *d_state, dpre))
        _lib.call("rssm_mtrssm_rollout_bwd", _mt_dims(actions, KL, KH, l_tau, h_tau, precision, obs_projected), w, inp, out, up, gin, gw)
        return [flat, d_actions, d_ea, d_ev, *d_state]  # flat = all weight grads, split by the caller


@mtrssm_rollout_bwd_op.register_fake
def _(weights, actions, embed_a, embed_v, state, feature, prior_h, prior_l, post_h, post_l, saved, d_feature, d_prior_h,  # noqa: ANN001
      d_prior_l, d_post_h, d_post_l, d_pz_h, d_pz_l, d_kl_l, d_kl_h, d_hidden_h, d_hidden_l, KL, KH, l_tau, h_tau, precision, kl_wq, kl_wp,
      obs_projected):
    return [actions.new_empty(sum(t.numel() for t in weights)), torch.empty_like(actions), torch.empty_like(embed_a),
            torch.empty_like(embed_v), *[torch.empty_like(t) for t in state]]


def _mt_setup(ctx, inputs, output) -> None:  # noqa: ANN001
    (weights, actions, embed_a, embed_v, state, u_post_l, u_post_h, u_prior_l, u_prior_h, KL, KH, l_tau, h_tau, precision,
     kl_wq, kl_wp, save, obs_projected) = inputs
    feature, hidden_h, hidden_l, prior_h, prior_l, post_h, post_l, pz_h, pz_l, kl_l, kl_h, saved = output
    if not save:
        raise RuntimeError("mtrssm_rollout was called with save=False but a gradient is required")
    ctx.set_materialize_grads(False)
    ctx.nw = len(weights)
    ctx.has_prior_stoch = u_prior_l is not None
    ctx.cfg = (KL, KH, l_tau, h_tau, precision, kl_wq, kl_wp, obs_projected)
    ctx.save_for_backward(*weights, actions, embed_a, embed_v, *state, feature, prior_h, prior_l, post_h, post_l, saved)


def _mt_backward(ctx, grads):  # noqa: ANN001
    d_feature, d_hh, d_hl, d_prior_h, d_prior_l, d_post_h, d_post_l, d_pz_h, d_pz_l, d_kl_l, d_kl_h, _ = grads
    t = ctx.saved_tensors
    weights = list(t[: ctx.nw])
    actions, embed_a, embed_v = t[ctx.nw: ctx.nw + 3]
    state = list(t[ctx.nw + 3: ctx.nw + 9])
    feature, prior_h, prior_l, post_h, post_l, saved = t[ctx.nw + 9:]
    KL, KH, l_tau, h_tau, precision, kl_wq, kl_wp, obs_projected = ctx.cfg
    hp = ctx.has_prior_stoch
    res = mtrssm_rollout_bwd_op(
        weights, actions, embed_a, embed_v, state, feature, prior_h, prior_l, post_h, post_l, saved, d_feature, d_prior_h,
        d_prior_l, d_post_h, d_post_l, d_pz_h if hp else None, d_pz_l if hp else None, d_kl_l, d_kl_h, d_hh, d_hl, KL, KH, l_tau, h_tau,
        precision, kl_wq, kl_wp, obs_projected,
    )
    flat, d_actions, d_ea, d_ev = res[:4]
    d_state = list(res[4:])
    gws = [g.view_as(w) for g, w in zip(flat.split([w.numel() for w in weights]), weights)]
    return (gws, d_actions, d_ea, d_ev, d_state, None, None, None, None, None, None, None, None, None, None, None, None, None)


mtrssm_rollout_op.register_autograd(_mt_backward, setup_context=_mt_setup)


# ---- bf16 fused policy: GROUPED outputs ----------------------------------------------------------------------------------------
# feature | hidden_h | hidden_l | the four probability tensors | the two prior draws of a (b,t) share ONE 1 KB row of a [B,T,256]
# buffer (kl_l | kl_h one [B,T,2]): a row-step writes eight whole 128-byte lines instead of ten partial-line segments (forward
# 0.559 -> 0.495 ms at the bench batch).  A custom op may not return tensors that alias each other, so the raw ops below return /
# take the row buffer, and `_MtrssmGroupedFn` (an autograd.Function over them) hands out the strided views and routes each view's
# gradient straight to the backward op -- autograd never assembles a dense gradient of the row.
def _fill_row_outputs(out, row: Tensor, kl: Tensor, has_prior: bool) -> None:  # noqa: ANN001
    base = row.data_ptr()
    for name, off in _lib.MT_ROW_OFFSETS.items():
        if name.startswith("prior_stoch") and not has_prior:
            continue
        setattr(out, name, base + 4 * off)
    out.kl_l, out.kl_h = kl.data_ptr(), kl.data_ptr() + 4
    out.ld_feature = out.ld_hidden = out.ld_probs = out.ld_stoch = _lib.MT_ROW_PITCH
    out.ld_kl = 2


@torch.library.custom_op("mtrssm_b200::mtrssm_rollout_grouped", mutates_args=())
def mtrssm_rollout_grouped_op(
    weights: Sequence[Tensor], actions: Tensor, embed_a: Tensor, embed_v: Tensor, state: Sequence[Tensor],
    u_post_l: Tensor, u_post_h: Tensor, u_prior_l: Optional[Tensor], u_prior_h: Optional[Tensor],
    KL: int, KH: int, l_tau: float, h_tau: float, precision: int, save: bool, obs_projected: bool,
) -> List[Tensor]:
    """-> [row [B,T,256], kl [B,T,2], saved]; no autograd of its own (see _MtrssmGroupedFn)."""
    with _on_device(actions, weights, embed_a, embed_v, state, u_post_l, u_post_h, u_prior_l, u_prior_h):
        _mt_check(weights, embed_a, state, obs_projected)
        B, T, _ = actions.shape
        dev = actions.device
        row = torch.empty(B, T, _lib.MT_ROW_PITCH, device=dev)
        kl = torch.empty(B, T, 2, device=dev)
        if u_prior_l is None:
            row[..., 224:].zero_()  # the prior-draw columns are not written
        saved = (torch.empty(_lib.mtrssm_saved_rows(B, precision), T, _lib.mtrssm_saved_elems(precision), device=dev,
                             dtype=_lib.record_dtype(precision)) if save else torch.empty(0, device=dev))
        w = _fill(_lib.MtrssmWeights(), _lib.MT_WEIGHT_FIELDS, weights)
        inp = _fill(_lib.MtrssmInputs(), ["actions", "embed_a", "embed_v", *_MT_STATE, "u_post_l", "u_post_h", "u_prior_l", "u_prior_h"],
                    (actions, embed_a, embed_v, *state, u_post_l, u_post_h, u_prior_l, u_prior_h))
        out = _lib.MtrssmOutputs()
        _fill_row_outputs(out, row, kl, u_prior_l is not None)
        out.saved = ptr(saved) if save else None
        _lib.call("rssm_mtrssm_rollout_fwd", _mt_dims(actions, KL, KH, l_tau, h_tau, precision, obs_projected), w, inp, out)
        return [row, kl, saved]


@mtrssm_rollout_grouped_op.register_fake
def _(weights, actions, embed_a, embed_v, state, u_post_l, u_post_h, u_prior_l, u_prior_h, KL, KH, l_tau, h_tau, precision, save,  # noqa: ANN001
      obs_projected):
    B, T, _ = actions.shape
    e = actions.new_empty
    return [e(B, T, _lib.MT_ROW_PITCH), e(B, T, 2),
            e(_lib.mtrssm_saved_rows(B, precision), T, _lib.mtrssm_saved_elems(precision), dtype=_lib.record_dtype(precision)) if save else e(0)]


@torch.library.custom_op("mtrssm_b200::mtrssm_rollout_grouped_bwd", mutates_args=())
def mtrssm_rollout_grouped_bwd_op(
    weights: Sequence[Tensor], actions: Tensor, embed_a: Tensor, embed_v: Tensor, state: Sequence[Tensor], row: Tensor, saved: Tensor,
    d_feature: Optional[Tensor], d_prior_h: Optional[Tensor], d_prior_l: Optional[Tensor], d_post_h: Optional[Tensor],
    d_post_l: Optional[Tensor], d_pz_h: Optional[Tensor], d_pz_l: Optional[Tensor], d_kl_l: Optional[Tensor], d_kl_h: Optional[Tensor],
    d_hidden_h: Optional[Tensor], d_hidden_l: Optional[Tensor],
    KL: int, KH: int, l_tau: float, h_tau: float, precision: int, kl_wq: float, kl_wp: float, obs_projected: bool,
) -> List[Tensor]:
    """Fused backward over the grouped row.  -> [flat weight grads, d_actions, d_embed_a, d_embed_v, *d_state]."""
    with _on_device(actions, weights, embed_a, embed_v, state, row, saved, d_feature, d_prior_h, d_prior_l, d_post_h, d_post_l, d_pz_h,
                    d_pz_l, d_kl_l, d_kl_h, d_hidden_h, d_hidden_l):
        B, T, A = actions.shape
        dev = actions.device
        if d_feature is None:
            d_feature = torch.zeros(B, T, 96, device=dev)
        d_hidden_h, d_hidden_l = _hidden_pair(d_hidden_h, d_hidden_l, B, T, dev)
        sizes = [t.numel() for t in weights]
        flat = torch.zeros(sum(sizes), device=dev)
        gws = [g.view_as(t) for g, t in zip(flat.split(sizes), weights)]
        e = lambda *s: torch.empty(*s, device=dev)  # noqa: E731
        EW = 32 if obs_projected else 64
        d_actions, d_ea, d_ev = e(B, T, A), e(B, T, EW), e(B, T, EW)
        d_state = [e(B, 32), e(B, 32), e(B, 32), e(B, 32), e(B, 16), e(B, 16)]
        w = _fill(_lib.MtrssmWeights(), _lib.MT_WEIGHT_FIELDS, weights)
        gw = _fill(_lib.MtrssmWeightGrads(), _lib.MT_WEIGHT_FIELDS, gws)
        inp = _fill(_lib.MtrssmInputs(), ["actions", "embed_a", "embed_v", *_MT_STATE], (actions, embed_a, embed_v, *state))
        out = _lib.MtrssmOutputs()
        _fill_row_outputs(out, row, row, True)  # the backward reads feature and the probabilities only
        out.kl_l = out.kl_h = None
        out.saved = ptr(saved)
        up = _fill(
            _lib.MtrssmUpstream(),
            ("d_feature d_prior_probs_h d_prior_probs_l d_post_probs_h d_post_probs_l d_prior_stoch_h d_prior_stoch_l d_kl_l d_kl_h "
             "d_hidden_h d_hidden_l").split(),
            tuple(_c(t) for t in (d_feature, d_prior_h, d_prior_l, d_post_h, d_post_l, d_pz_h, d_pz_l, d_kl_l, d_kl_h, d_hidden_h,
                                  d_hidden_l)),
        )
        up.kl_wq, up.kl_wp = kl_wq, kl_wp
        gin = _fill(_lib.MtrssmInputGrads(), ["d_actions", "d_embed_a", "d_embed_v", *("d_" + n for n in _MT_STATE)],
                    (d_actions, d_ea, d_ev, *d_state))
        _lib.call("rssm_mtrssm_rollout_bwd", _mt_dims(actions, KL, KH, l_tau, h_tau, precision, obs_projected), w, inp, out, up, gin, gw)
        return [flat, d_actions, d_ea, d_ev, *d_state]


@mtrssm_rollout_grouped_bwd_op.register_fake
def _(weights, actions, embed_a, embed_v, state, row, saved, d_feature, d_prior_h, d_prior_l, d_post_h, d_post_l, d_pz_h, d_pz_l,  # noqa: ANN001
      d_kl_l, d_kl_h, d_hidden_h, d_hidden_l, KL, KH, l_tau, h_tau, precision, kl_wq, kl_wp, obs_projected):
    return [actions.new_empty(sum(t.numel() for t in weights)), torch.empty_like(actions), torch.empty_like(embed_a),
            torch.empty_like(embed_v), *[torch.empty_like(t) for t in state]]


class _MtrssmGroupedFn(torch.autograd.Function):
    """Autograd of the grouped rollout.  forward(cfg, *tensors): tensors = 28 weights, actions, embed_a, embed_v, 6 state tensors,
    u_post_l, u_post_h, u_prior_l, u_prior_h (the last two may be None) -> the 11 outputs of `mtrssm_rollout` as strided views."""

    NW = len(_lib.MT_WEIGHT_FIELDS)

    @staticmethod
    def forward(ctx, cfg, *t):  # noqa: ANN001
        nw = _MtrssmGroupedFn.NW
        weights, (actions, embed_a, embed_v), state = list(t[:nw]), t[nw:nw + 3], list(t[nw + 3:nw + 9])
        u_post_l, u_post_h, u_prior_l, u_prior_h = t[nw + 9:nw + 13]
        KL, KH, l_tau, h_tau, precision, kl_wq, kl_wp, save, obs_projected = cfg
        row, kl, saved = mtrssm_rollout_grouped_op(weights, actions, embed_a, embed_v, state, u_post_l, u_post_h, u_prior_l, u_prior_h,
                                                   KL, KH, l_tau, h_tau, precision, save, obs_projected)
        o = _lib.MT_ROW_OFFSETS
        B, T, _ = actions.shape
        v = lambda name, w: row[..., o[name]:o[name] + w]  # noqa: E731
        has_prior = u_prior_l is not None
        outs = (
            v("feature", 96), v("hidden_h", 32), v("hidden_l", 32),
            v("prior_probs_h", 16).unflatten(-1, (16 // KH, KH)), v("prior_probs_l", 16).unflatten(-1, (16 // KL, KL)),
            v("post_probs_h", 16).unflatten(-1, (16 // KH, KH)), v("post_probs_l", 16).unflatten(-1, (16 // KL, KL)),
            v("prior_stoch_h", 16) if has_prior else None, v("prior_stoch_l", 16) if has_prior else None,
            kl[..., 0], kl[..., 1],
        )
        ctx.set_materialize_grads(False)
        if save:
            ctx.cfg, ctx.has_prior = cfg, has_prior
            ctx.save_for_backward(*weights, actions, embed_a, embed_v, *state, row, saved)
        return outs

    @staticmethod
    def backward(ctx, *g):  # noqa: ANN001
        d_feature, d_hh, d_hl, d_prior_h, d_prior_l, d_post_h, d_post_l, d_pz_h, d_pz_l, d_kl_l, d_kl_h = g
        nw = _MtrssmGroupedFn.NW
        t = ctx.saved_tensors
        weights, (actions, embed_a, embed_v), state = list(t[:nw]), t[nw:nw + 3], list(t[nw + 3:nw + 9])
        row, saved = t[nw + 9:]
        KL, KH, l_tau, h_tau, precision, kl_wq, kl_wp, _save, obs_projected = ctx.cfg
        hp = ctx.has_prior
        res = mtrssm_rollout_grouped_bwd_op(
            weights, actions, embed_a, embed_v, state, row, saved, d_feature, d_prior_h, d_prior_l, d_post_h, d_post_l,
            d_pz_h if hp else None, d_pz_l if hp else None, d_kl_l, d_kl_h, d_hh, d_hl, KL, KH, l_tau, h_tau, precision, kl_wq, kl_wp,
            obs_projected,
        )
        flat, d_actions, d_ea, d_ev = res[:4]
        gws = [x.view_as(w) for x, w in zip(flat.split([w.numel() for w in weights]), weights)]
        return (None, *gws, d_actions, d_ea, d_ev, *res[4:], None, None, None, None)


def mtrssm_rollout(
    weights: Sequence[Tensor], *, actions: Tensor, embed_a: Tensor, embed_v: Tensor, deter_h0: Tensor, deter_l0: Tensor,
    hidden_h0: Tensor, hidden_l0: Tensor, stoch_h0: Tensor, stoch_l0: Tensor, u_post_l: Tensor, u_post_h: Tensor,
    u_prior_l: Optional[Tensor] = None, u_prior_h: Optional[Tensor] = None, class_size_l: int = 4, class_size_h: int = 2,
    l_tau: float = 2.0, h_tau: float = 4.0, precision: int = _lib.PRECISION_FP32, use_kl_balancing: bool = True,
    obs_projected: bool = False,
) -> dict[str, Tensor]:
    """Fused MoPoE-MMTRSSM rollout_representation on encoder outputs (mmtrssm/mopoe_mmtrssm/core.py:364-494).

    feature [B,T,96] = [deter_h | stoch_h | deter_l | stoch_l] (mmtrssm/state.py:51).

    `obs_projected=True` (bf16 fused policy only): `embed_a` / `embed_v` are the PRE-MULTIPLIED first-layer
    partials `e @ W1[:, 32:].T` of the two modality heads, [B,T,32] -- one big GEMM before the loop (or the encoder's last Linear
    with the merged weight) instead of 2 x 16 small MMAs and 512 B of embedding reads per (b,t) inside it.  The kernels use
    `W1[:, :32]` only; the returned gradients of `embed_*` are those of the partials, and `W1[:, 32:]` gets its gradient from the
    caller's GEMM through autograd (`obs_projection` below does exactly that).
    """
    state = [deter_h0, deter_l0, hidden_h0, hidden_l0, stoch_h0, stoch_l0]
    save = torch.is_grad_enabled() and any(t.requires_grad for t in (*weights, actions, embed_a, embed_v, *state))
    wq, wp = kl_path_weights(use_kl_balancing)
    names = ("feature hidden_h hidden_l prior_probs_h prior_probs_l post_probs_h post_probs_l prior_stoch_h prior_stoch_l "
             "kl_l kl_h").split()
    if precision == _lib.PRECISION_BF16_FUSED:  # grouped outputs (one 1 KB row per (b,t)), tile-blocked record, fused backward
        cfg = (class_size_l, class_size_h, float(l_tau), float(h_tau), precision, wq, wp, save, bool(obs_projected))
        outs = _MtrssmGroupedFn.apply(cfg, *[_c(w) for w in weights], _c(actions), _c(embed_a), _c(embed_v), *[_c(s) for s in state],
                                      _c(u_post_l), _c(u_post_h), _c(u_prior_l), _c(u_prior_h))
        return dict(zip(names, outs))
    out = mtrssm_rollout_op(
        [_c(w) for w in weights], _c(actions), _c(embed_a), _c(embed_v), [_c(s) for s in state], _c(u_post_l), _c(u_post_h),
        _c(u_prior_l), _c(u_prior_h),
        class_size_l, class_size_h, float(l_tau), float(h_tau), precision, wq, wp, save, obs_projected,
    )
    res = dict(zip(names, out[:-1]))
    if u_prior_l is None:
        res["prior_stoch_h"] = res["prior_stoch_l"] = None
    return res


def obs_projection(embed: Tensor, w1: Tensor, deter_size: int = 32) -> Tensor:
    """The hoisted half of a modality head's first layer: `embed @ W1[:, deter_size:].T` for ALL (b,t) in one GEMM
    (fp32, differentiable: autograd gives `d embed` and the `W1[:, deter_size:]` columns of `d W1`).  Feed the result to
    `mtrssm_rollout(..., obs_projected=True)`."""
    return torch.nn.functional.linear(embed.float(), w1[:, deter_size:].float())


@torch.library.custom_op("mtrssm_b200::mtrssm_imagine", mutates_args=())
def mtrssm_imagine_op(weights: Sequence[Tensor], actions: Tensor, state: Sequence[Tensor], u_l: Tensor, u_h: Tensor,
                      KL: int, KH: int, l_tau: float, h_tau: float, precision: int) -> List[Tensor]:
    with _on_device(actions, weights, state, u_l, u_h):
        B, T, _ = actions.shape
        dev = actions.device
        e = lambda *s: torch.empty(*s, device=dev)  # noqa: E731
        feature, hidden_h, hidden_l = e(B, T, 96), e(B, T, 32), e(B, T, 32)
        probs_h, probs_l = e(B, T, 16 // KH, KH), e(B, T, 16 // KL, KL)
        w = _fill(_lib.MtrssmWeights(), _lib.MT_WEIGHT_FIELDS, weights)
        inp = _fill(_lib.MtrssmInputs(), ["actions", *_MT_STATE, "u_prior_l", "u_prior_h"], (actions, *state, u_l, u_h))
        out = _fill(_lib.MtrssmOutputs(), "feature hidden_h hidden_l prior_probs_h prior_probs_l".split(),
                    (feature, hidden_h, hidden_l, probs_h, probs_l))
        _lib.call("rssm_mtrssm_imagine_fwd", _mt_dims(actions, KL, KH, l_tau, h_tau, precision), w, inp, out)
        return [feature, hidden_h, hidden_l, probs_h, probs_l]


@mtrssm_imagine_op.register_fake
def _(weights, actions, state, u_l, u_h, KL, KH, l_tau, h_tau, precision):  # noqa: ANN001
    B, T, _ = actions.shape
    e = actions.new_empty
    return [e(B, T, 96), e(B, T, 32), e(B, T, 32), e(B, T, 16 // KH, KH), e(B, T, 16 // KL, KL)]


def mtrssm_imagine(
    weights: Sequence[Tensor], *, actions: Tensor, deter_h0: Tensor, deter_l0: Tensor, hidden_h0: Tensor, hidden_l0: Tensor,
    stoch_h0: Tensor, stoch_l0: Tensor, u_l: Tensor, u_h: Tensor, class_size_l: int = 4, class_size_h: int = 2,
    l_tau: float = 2.0, h_tau: float = 4.0, precision: int = _lib.PRECISION_FP32,
) -> dict[str, Tensor]:
    """Fused MoPoE_MMTRSSM.rollout_transition (mmtrssm/mopoe_mmtrssm/core.py:496-544).  Forward only."""
    state = [_c(s) for s in (deter_h0, deter_l0, hidden_h0, hidden_l0, stoch_h0, stoch_l0)]
    with torch.no_grad():
        out = mtrssm_imagine_op([_c(w) for w in weights], _c(actions), state, _c(u_l), _c(u_h), class_size_l, class_size_h,
                                float(l_tau), float(h_tau), precision)
    return dict(zip("feature hidden_h hidden_l probs_h probs_l".split(), out))
