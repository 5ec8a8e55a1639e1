"""ctypes binding of librssm_rollout.so (C ABI: include/rssm_rollout.h).

There is NO fallback: if the library is missing it is built in-tree with nvcc; if that fails, or a
call returns non-zero, a RuntimeError is raised.
"""

from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache

import torch

from .build import LIB, build_library

PRECISION_FP32 = 0
PRECISION_BF16 = 1
PRECISION_BF16_FUSED = 2  # bf16 path with the one-kernel backward (tcgen05 weight gradients); MMTRSSM only
ABI_VERSION = 5
MRSSM_SAVED_FLOATS, MRSSM_DPRE_FLOATS = 320, 336
MTRSSM_SAVED_FLOATS, MTRSSM_DPRE_FLOATS = 208, 304
MTRSSM_SAVED_BF16 = MTRSSM_SAVED_FLOATS

_fp = C.c_void_p  # device pointers travel as integers


def _struct(name: str, fields: list[tuple[str, type]]) -> type:
    return type(name, (C.Structure,), {"_fields_": fields})


def _ptrs(names: str) -> list[tuple[str, type]]:
    return [(n, _fp) for n in names.split()]


_MR_W = "asp_w1 asp_b1 asp_w2 asp_b2 w_ih w_hh b_ih b_hh pr_w1 pr_b1 pr_w2 pr_b2 au_w1 au_b1 au_w2 au_b2 vi_w1 vi_b1 vi_w2 vi_b2"
_MT_W = (
    "l_d2h_w l_d2h_b l_in_w l_in_b h_d2h_w h_d2h_b h_in_w h_in_b lp_w1 lp_b1 lp_w2 lp_b2 hp_w1 hp_b1 hp_w2 hp_b2 "
    "hq_w1 hq_b1 hq_w2 hq_b2 au_w1 au_b1 au_w2 au_b2 vi_w1 vi_b1 vi_w2 vi_b2"
)
MR_WEIGHT_FIELDS = tuple(_MR_W.split())
MT_WEIGHT_FIELDS = tuple(_MT_W.split())

MrssmDims = _struct("RssmMrssmDims", [(n, C.c_int) for n in "B T A E D H C K precision unimodal".split()])
MrssmWeights = _struct("RssmMrssmWeights", _ptrs(_MR_W))
MrssmWeightGrads = _struct("RssmMrssmWeightGrads", _ptrs(_MR_W))
MrssmInputs = _struct("RssmMrssmInputs", _ptrs("actions embed_a embed_v h0 z0 u_post u_prior"))
MrssmOutputs = _struct(
    "RssmMrssmOutputs", _ptrs("feature prior_probs post_probs prior_stoch kl saved workspace") + [("workspace_bytes", C.c_size_t)]
)
MrssmUpstream = _struct(
    "RssmMrssmUpstream",
    _ptrs("d_feature d_prior_probs d_post_probs d_prior_stoch d_kl") + [("kl_wq", C.c_float), ("kl_wp", C.c_float)],
)
MrssmInputGrads = _struct(
    "RssmMrssmInputGrads", _ptrs("d_actions d_embed_a d_embed_v d_h0 d_z0 dpre workspace") + [("workspace_bytes", C.c_size_t)]
)

MtrssmDims = _struct(
    "RssmMtrssmDims",
    [(n, C.c_int) for n in "B T A E HD LD HH HR CL KL CH KH".split()]
    + [("l_tau", C.c_float), ("h_tau", C.c_float), ("precision", C.c_int), ("obs_projected", C.c_int)],
)
MtrssmWeights = _struct("RssmMtrssmWeights", _ptrs(_MT_W))
MtrssmWeightGrads = _struct("RssmMtrssmWeightGrads", _ptrs(_MT_W))
MtrssmInputs = _struct(
    "RssmMtrssmInputs",
    _ptrs(
        "actions embed_a embed_v deter_h0 deter_l0 hidden_h0 hidden_l0 stoch_h0 stoch_l0 "
        "u_post_l u_post_h u_prior_l u_prior_h"
    ),
)
MtrssmOutputs = _struct(
    "RssmMtrssmOutputs",
    _ptrs(
        "feature hidden_h hidden_l prior_probs_h prior_probs_l post_probs_h post_probs_l "
        "prior_stoch_h prior_stoch_l kl_l kl_h saved"
    )
    + [(n, C.c_int) for n in "ld_feature ld_hidden ld_probs ld_stoch ld_kl".split()],  # row pitches, 0 = natural (ABI v5)
)
# grouped per-(b,t) output row of the bf16 fused policy (include/rssm_rollout.h): float offsets inside the [B,T,256] buffer
MT_ROW_PITCH = 256
MT_ROW_OFFSETS = {"feature": 0, "hidden_h": 96, "hidden_l": 128, "prior_probs_h": 160, "prior_probs_l": 176, "post_probs_h": 192,
                  "post_probs_l": 208, "prior_stoch_h": 224, "prior_stoch_l": 240}
MtrssmUpstream = _struct(
    "RssmMtrssmUpstream",
    _ptrs(
        "d_feature d_prior_probs_h d_prior_probs_l d_post_probs_h d_post_probs_l d_prior_stoch_h d_prior_stoch_l "
        "d_kl_l d_kl_h"
    )
    + [("kl_wq", C.c_float), ("kl_wp", C.c_float)]
    + _ptrs("d_hidden_h d_hidden_l"),
)
MtrssmInputGrads = _struct(
    "RssmMtrssmInputGrads",
    _ptrs("d_actions d_embed_a d_embed_v d_deter_h0 d_deter_l0 d_hidden_h0 d_hidden_l0 d_stoch_h0 d_stoch_l0 dpre"),
)

NLL_MAX_SEGMENTS = 4
DTYPE_CODES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
NllPair = _struct(
    "RssmNllPair",
    [("prediction", _fp), ("target", _fp), ("n_elems", C.c_size_t), ("n_batch", C.c_size_t), ("scale", C.c_float),
     ("loss", _fp), ("d_loss", _fp), ("d_prediction", _fp), ("d_target", _fp)],
)

EXPORTS = (
    "rssm_mrssm_rollout_fwd", "rssm_mrssm_rollout_bwd", "rssm_mrssm_imagine_fwd",
    "rssm_mtrssm_rollout_fwd", "rssm_mtrssm_rollout_bwd", "rssm_mtrssm_imagine_fwd",
    "rssm_mrssm_wgrad", "rssm_mtrssm_wgrad",
    "rssm_abi_version", "rssm_last_error", "rssm_kernel_launch_count",
    "rssm_mrssm_saved_bytes", "rssm_mrssm_workspace_bytes",
    "rssm_gaussian_nll_fwd", "rssm_gaussian_nll_bwd", "rssm_gaussian_nll_workspace_bytes",
    "rssm_p2p_region_bytes", "rssm_p2p_alloc", "rssm_p2p_free", "rssm_p2p_export", "rssm_p2p_import", "rssm_p2p_close",
    "rssm_p2p_bucket", "rssm_p2p_allreduce_mean", "rssm_p2p_status",
)
P2P_MAX_RANKS = 8
P2pComm = _struct("RssmP2pComm", [("world", C.c_int), ("rank", C.c_int), ("regions", C.c_void_p * P2P_MAX_RANKS), ("n", C.c_size_t)])


@lru_cache(maxsize=1)
def lib() -> C.CDLL:
    """Load (building first if needed) the CUDA library.  Raises if it cannot be had."""
    override = os.environ.get("RSSM_ROLLOUT_LIB")  # experiments / profiling builds of the same ABI
    path = override if override else build_library()
    handle = C.CDLL(str(path))
    for name in EXPORTS:
        if not hasattr(handle, name):
            raise RuntimeError(f"{LIB.name} does not export {name}")
    handle.rssm_last_error.restype = C.c_char_p
    handle.rssm_kernel_launch_count.restype = C.c_longlong
    P = C.c_void_p
    for name in EXPORTS[:8]:
        getattr(handle, name).restype = C.c_int
    handle.rssm_mrssm_saved_bytes.restype = C.c_size_t
    handle.rssm_mrssm_saved_bytes.argtypes = [P]
    handle.rssm_mrssm_workspace_bytes.restype = C.c_size_t
    handle.rssm_mrssm_workspace_bytes.argtypes = [P, C.c_int]
    handle.rssm_mrssm_wgrad.argtypes = [P] * 6
    handle.rssm_mtrssm_wgrad.argtypes = [P] * 6
    handle.rssm_mrssm_rollout_fwd.argtypes = [P] * 5
    handle.rssm_mrssm_imagine_fwd.argtypes = [P] * 5
    handle.rssm_mrssm_rollout_bwd.argtypes = [P] * 8
    handle.rssm_mtrssm_rollout_fwd.argtypes = [P] * 5
    handle.rssm_mtrssm_imagine_fwd.argtypes = [P] * 5
    handle.rssm_mtrssm_rollout_bwd.argtypes = [P] * 8
    handle.rssm_gaussian_nll_workspace_bytes.restype = C.c_size_t
    handle.rssm_gaussian_nll_workspace_bytes.argtypes = []
    handle.rssm_gaussian_nll_fwd.restype = C.c_int
    handle.rssm_gaussian_nll_fwd.argtypes = [P, C.c_int, C.c_int, P, C.c_size_t, P]
    handle.rssm_gaussian_nll_bwd.restype = C.c_int
    handle.rssm_gaussian_nll_bwd.argtypes = [P, C.c_int, C.c_int, P]
    handle.rssm_p2p_region_bytes.restype = C.c_size_t
    handle.rssm_p2p_region_bytes.argtypes = [C.c_size_t]
    handle.rssm_p2p_alloc.argtypes = [C.c_size_t, C.POINTER(P)]
    handle.rssm_p2p_free.argtypes = [P]
    handle.rssm_p2p_export.argtypes = [P, C.c_char_p]
    handle.rssm_p2p_import.argtypes = [C.c_char_p, C.POINTER(P)]
    handle.rssm_p2p_close.argtypes = [P]
    handle.rssm_p2p_bucket.restype = P
    handle.rssm_p2p_bucket.argtypes = [P, C.c_size_t, C.c_int]
    handle.rssm_p2p_allreduce_mean.argtypes = [P, C.c_longlong, P, P, C.c_int, P]
    handle.rssm_p2p_status.argtypes = [P]
    if handle.rssm_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{LIB.name}: ABI version {handle.rssm_abi_version()} != {ABI_VERSION}")
    return handle


def call(fn_name: str, *args) -> None:  # noqa: ANN002
    """Invoke an entry point on the current CUDA stream; non-zero status -> RuntimeError."""
    handle = lib()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    cargs = [None if a is None else (C.c_void_p(a) if isinstance(a, int) else C.byref(a)) for a in args]
    status = getattr(handle, fn_name)(*cargs, stream)
    if status != 0:
        raise RuntimeError(f"{fn_name} failed: {handle.rssm_last_error().decode()}")


def ptr(t: torch.Tensor | None) -> int | None:
    """Device pointer of a contiguous fp32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not (t.is_cuda and t.dtype in (torch.float32, torch.bfloat16) and t.is_contiguous()):
        raise RuntimeError(
            f"rollout kernels need contiguous fp32 CUDA tensors (bf16 only for the opaque records), got device={t.device} "
            f"dtype={t.dtype} contiguous={t.is_contiguous()}"
        )
    return t.data_ptr()


def record_dtype(precision: int) -> torch.dtype:
    """Element type of the opaque saved / dpre records (include/rssm_rollout.h)."""
    return torch.float32 if precision == PRECISION_FP32 else torch.bfloat16


def mtrssm_saved_rows(B: int, precision: int) -> int:
    """Rows of the MMTRSSM saved record: B, or B rounded up to the 16-sequence tile for the tile-blocked record of
    PRECISION_BF16_FUSED (include/rssm_rollout.h, RssmMtrssmOutputs.saved)."""
    return (B + 15) // 16 * 16 if precision == PRECISION_BF16_FUSED else B


def mtrssm_saved_elems(precision: int) -> int:
    """Elements per (b,t) of the MMTRSSM saved record (include/rssm_rollout.h)."""
    return MTRSSM_SAVED_FLOATS


def mrssm_saved_bytes(dims) -> int:  # noqa: ANN001
    """Bytes of the opaque MRSSM saved record for `dims` (0 = unsupported sizes; the launch reports the error)."""
    return int(lib().rssm_mrssm_saved_bytes(C.byref(dims)))


def mrssm_workspace_bytes(dims, backward: bool) -> int:  # noqa: ANN001
    """Bytes of scratch a wide-family MRSSM call needs (0 for the default family)."""
    return int(lib().rssm_mrssm_workspace_bytes(C.byref(dims), 1 if backward else 0))


def launch_count() -> int:
    return int(lib().rssm_kernel_launch_count())
