"""GPU parity AT THE BENCHMARKED CONFIGURATIONS (BASELINE.json configs[0..3]): the CUDA path (custom op -> ctypes -> C ABI) against
the oracle, which at these sizes runs eager in fp32 ON THE GPU (`oracle/rssm_oracle.py` is device-agnostic torch; TF32 is off, so
its matmuls are true fp32) -- the CPU would need minutes per case.

What is compared: a rollout of the reduced-precision path cannot be compared draw-for-draw with an fp32 rollout (one flipped
categorical draw changes the whole future), so -- exactly as in tests/test_rollout_gpu.py -- the oracle is TEACHER-FORCED on the
kernel's own posterior draws and the kernel's draws are separately checked to be the inverse-CDF draws of ITS OWN probabilities.
Every state, probability and KL term, and every gradient (inputs, initial state, all weights) is compared.

STATED bf16 TOLERANCES (the `north_star` "stated bf16 tolerance for the tensor-core path"), asserted below:

* states / probabilities (O(1) quantities): the bf16 error does NOT grow with the horizon -- the leaky integrators (tau = 2, 4), the
  GRU gates and tanh are contractions and teacher-forcing removes the discrete divergence -- so the bound is FLAT IN T:
      max |err| <= STATE_MAX (4e-2 default family, 6e-2 hidden 512)   and   rms err <= STATE_RMS (6e-3)
  for every t up to T = 512 (cfg4); the per-step curve is recorded in `gpurun_out/parity_r2.json` and asserted step by step
  (measured, `profiles/r2_a_parity.json`: worst step of T = 512 0.0295, worst 64-step window maxima 0.023 .. 0.029 with no
  trend; rms 0.0039; B = 37888, T = 30: 0.0275 / 0.0032; hidden 512, T = 64: 0.0046 / 0.0003).
* KL per (b,t): |err| <= 5e-2 + 5e-2 |kl|.
* gradients: max |err| <= GRAD_MAX (2e-2; hidden 512: 3e-2) of the tensor's scale (max |grad|) and rms err <= GRAD_RMS (1.5e-2) of
  the tensor's rms -- for every T tested (8 .. 512), batch up to the bench batch (measured: max 0.017 of scale, rms 0.0087 of rms
  at T = 512 with the losses' outputs receiving gradients, 0.0128 when the MTRNN.hidden outputs receive random gradients too, as
  the tests below do; 0.014 / 0.0060 at the bench batch; 0.0059 / 0.0047 at hidden 512).
The fp32-parity path (precision 0) is compared at the same large sizes with the module tolerances (1e-5 / 1e-4).
"""

from __future__ import annotations

import json
import os
from pathlib import Path

import pytest
import torch

from oracle import rssm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

STATE_MAX, STATE_RMS = 4e-2, 6e-3
GRAD_MAX, GRAD_RMS = 2e-2, 1.5e-2
WIDE_STATE_MAX, WIDE_STATE_RMS, WIDE_GRAD_MAX, WIDE_GRAD_RMS = 6e-2, 6e-3, 3e-2, 1.2e-2

_LOG: dict = {}
_LOG_PATH = Path(os.environ.get("GRAFT_REPO_ROOT", Path(__file__).resolve().parent.parent)) / "gpurun_out" / "parity_r2.json"


def _flush_log() -> None:
    try:
        _LOG_PATH.parent.mkdir(exist_ok=True)
        _LOG_PATH.write_text(json.dumps(_LOG, indent=1))
    except OSError:
        pass


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from multimodal_mtrssm_b200 import params as P
    from multimodal_mtrssm_b200 import rollout_ops as R

    return R, P


def cuda(d):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


class Stats:
    """max-abs and rms error per tensor against a stated (max, rms) bound; logs everything, asserts at the end."""

    def __init__(self, title: str) -> None:
        self.title, self.rows, self.failed = title, {}, []

    def state(self, name, got, want, tol_max, tol_rms):
        err = (got.detach().float() - want.detach().float())
        m, r = float(err.abs().max()), float(err.pow(2).mean().sqrt())
        ok = m <= tol_max and r <= tol_rms and bool(torch.isfinite(got).all())
        self.rows[name] = dict(max_abs=m, rms=r, tol_max=tol_max, tol_rms=tol_rms, ok=ok)
        if not ok:
            self.failed.append(name)

    def grad(self, name, got, want, tol_max, tol_rms):
        got, want = got.detach().float(), want.detach().float()
        err = got - want
        scale, rms = max(float(want.abs().max()), 1e-6), max(float(want.pow(2).mean().sqrt()), 1e-7)
        m, r = float(err.abs().max()) / scale, float(err.pow(2).mean().sqrt()) / rms
        ok = m <= tol_max and r <= tol_rms and bool(torch.isfinite(got).all())
        self.rows[name] = dict(max_of_scale=m, rms_of_rms=r, scale=scale, tol_max=tol_max, tol_rms=tol_rms, ok=ok)
        if not ok:
            self.failed.append(name)

    def finish(self):
        _LOG[self.title] = self.rows
        _flush_log()
        print(f"\n== {self.title}")
        for k, v in self.rows.items():
            print(f"{k:44s} " + " ".join(f"{a}={b:.3e}" if isinstance(b, float) else f"{a}={b}" for a, b in v.items()))
        assert not self.failed, f"{self.title}: out of the stated tolerance: {self.failed}"


# =====================================================================================================================
# MoPoE-MMTRSSM, default.yaml dims (cfg2 / cfg4 / cfg5; the headline bench runs B = 37888, T = 30, bf16 fused backward)
# =====================================================================================================================
MT_GRAD_IN = ("actions", "embed_a", "embed_v", "deter_h0", "deter_l0", "hidden_h0", "hidden_l0", "stoch_h0", "stoch_l0")
MT_STATE_KEYS = ("hidden_h", "hidden_l", "prior_probs_h", "prior_probs_l", "post_probs_h", "post_probs_l")


def mt_upstream(B, T, dims, bench_loss: bool, hidden: bool = True):
    """bench_loss: exactly what bench.py backpropagates (d_feature ~ N(0,1), d_kl = 1); otherwise every output gets a gradient
    (`hidden`: including the MTRNN.hidden outputs)."""
    g = torch.Generator(device="cuda").manual_seed(7)
    r = lambda *s: torch.randn(*s, generator=g, device="cuda")  # noqa: E731
    up = {"feature": r(B, T, 96), "kl_l": torch.ones(B, T, device="cuda"), "kl_h": torch.ones(B, T, device="cuda")}
    if not bench_loss:
        up.update(kl_l=r(B, T), kl_h=r(B, T), post_probs_l=r(B, T, dims["CL"], dims["KL"]), post_probs_h=r(B, T, dims["CH"], dims["KH"]),
                  prior_probs_l=r(B, T, dims["CL"], dims["KL"]), prior_probs_h=r(B, T, dims["CH"], dims["KH"]))
        if hidden:
            up.update(hidden_h=r(B, T, 32), hidden_l=r(B, T, 32))
    return up


def run_mt_both(R, P, B, T, precision, bench_loss, dims=H.MT_DIMS, gain=2.0):
    """kernel (given precision) and GPU-eager fp32 oracle teacher-forced on the kernel's draws; returns everything to compare"""
    params = H.make_params(H.MT_SHAPES, gain=gain)
    inp = H.mtrssm_inputs(B, T, dims)
    inp["u_prior_l"] = inp["u_prior_h"] = None  # the prior's own draws are not part of the training loss
    up = mt_upstream(B, T, dims, bench_loss)
    w = {k: v.cuda().requires_grad_(True) for k, v in params.items()}
    x = cuda(inp)
    for k in MT_GRAD_IN:
        x[k] = x[k].requires_grad_(True)
    got = R.mtrssm_rollout(P.mtrssm_weight_list(w), class_size_l=dims["KL"], class_size_h=dims["KH"], l_tau=dims["l_tau"],
                           h_tau=dims["h_tau"], precision=precision, **x)
    sum((got[k] * up[k]).sum() for k in up).backward()
    f = got["feature"].detach()
    idx_h = f[..., 32:48].reshape(B, T, dims["CH"], dims["KH"]).argmax(-1)
    idx_l = f[..., 80:].reshape(B, T, dims["CL"], dims["KL"]).argmax(-1)
    w_ref = {k: v.cuda().requires_grad_(True) for k, v in params.items()}
    x_ref = cuda(inp)
    for k in MT_GRAD_IN:
        x_ref[k] = x_ref[k].requires_grad_(True)
    want = O.mtrssm_rollout(w_ref, dims=dims, forced_idx_l=idx_l, forced_idx_h=idx_h, **x_ref)
    want["kl_l"] = O.kl_per_sample(want["post_probs_l"], want["prior_probs_l"], True)
    want["kl_h"] = O.kl_per_sample(want["post_probs_h"], want["prior_probs_h"], True)
    want["feature"] = want["post_feature"]
    sum((want[k] * up[k]).sum() for k in up).backward()
    torch.cuda.synchronize()
    return got, w, x, want, w_ref, x_ref, inp


def check_mt(st: Stats, got, w, x, want, w_ref, x_ref, inp, dims, smax, srms, gmax, grms):
    f = got["feature"]
    st.state("deter_h", f[..., :32], want["deter_h"], smax, srms)
    st.state("deter_l", f[..., 48:80], want["deter_l"], smax, srms)
    for k in MT_STATE_KEYS:
        st.state(k, got[k], want[k], smax, srms)
    for k in ("kl_l", "kl_h"):
        err = (got[k] - want[k]).abs()
        bound = 5e-2 + 5e-2 * want[k].abs()
        st.rows[k] = dict(max_abs=float(err.max()), ok=bool((err <= bound).all()))
        if not st.rows[k]["ok"]:
            st.failed.append(k)
    # the kernel's draws are the inverse-CDF draws of its own probabilities, exact one-hots
    for probs, u, sl, C, K in ((got["post_probs_l"], inp["u_post_l"], slice(80, 96), dims["CL"], dims["KL"]),
                               (got["post_probs_h"], inp["u_post_h"], slice(32, 48), dims["CH"], dims["KH"])):
        z = f[..., sl].reshape(*f.shape[:2], C, K)
        assert bool(((z == 0) | (z == 1)).all()) and bool((z.sum(-1) == 1).all())
        u = u.cuda()
        self_idx, margin = O.inverse_cdf_index(probs.detach(), u), O.cdf_margin(probs.detach(), u)
        assert bool(((self_idx == z.argmax(-1)) | (margin < 1e-5)).all())
    for k in MT_GRAD_IN:
        st.grad("d " + k, x[k].grad, x_ref[k].grad, gmax, grms)
    for k in w:
        st.grad("d " + k.replace("rnn_to_post_projector", "post"), w[k].grad, w_ref[k].grad, gmax, grms)


@pytest.mark.parametrize("B,bench_loss", [(4096, False), (37888, True)])
def test_cfg2_bf16_fused_at_bench_size_vs_oracle(ops, B, bench_loss):
    """The HEADLINE configuration (MoPoE-MMTRSSM default dims, T = 30, bf16 operands, tcgen05-fused backward): B = 4096 with a
    gradient on every output, and bench.py's own batch (37888 = 148 SMs x 256) with bench.py's own loss."""
    R, P = ops
    from multimodal_mtrssm_b200 import _lib

    res = run_mt_both(R, P, B, 30, _lib.PRECISION_BF16_FUSED, bench_loss)
    st = Stats(f"cfg2 MMTRSSM bf16-fused B={B} T=30 vs GPU-eager fp32 oracle (teacher-forced)")
    check_mt(st, *res, H.MT_DIMS, STATE_MAX, STATE_RMS, GRAD_MAX, GRAD_RMS)
    st.finish()


def test_cfg2_fp32_path_at_bench_size_vs_oracle(ops):
    """The fp32-parity policy at B = 4096, T = 30: module tolerances (1e-5 relative states, 1e-4 gradients), draws NOT forced --
    uniforms near a CDF boundary of the oracle trajectory are excluded by checking the draws agree wherever the margin is > 2e-4."""
    R, P = ops
    B, T, dims = 4096, 30, H.MT_DIMS
    params = H.make_params(H.MT_SHAPES)
    inp = H.mtrssm_inputs(B, T, dims)
    inp["u_prior_l"] = inp["u_prior_h"] = None
    w = {k: v.cuda().requires_grad_(True) for k, v in params.items()}
    x = cuda(inp)
    for k in MT_GRAD_IN:
        x[k] = x[k].requires_grad_(True)
    got = R.mtrssm_rollout(P.mtrssm_weight_list(w), precision=0, **x)
    f = got["feature"].detach()
    idx_h = f[..., 32:48].reshape(B, T, 8, 2).argmax(-1)
    idx_l = f[..., 80:].reshape(B, T, 4, 4).argmax(-1)
    w_ref = {k: v.cuda().requires_grad_(True) for k, v in params.items()}
    x_ref = cuda(inp)
    for k in MT_GRAD_IN:
        x_ref[k] = x_ref[k].requires_grad_(True)
    want = O.mtrssm_rollout(w_ref, dims=dims, forced_idx_l=idx_l, forced_idx_h=idx_h, **x_ref)
    # where the oracle's own (unforced) draw is not on a knife edge it must equal the kernel's draw
    own_l = O.inverse_cdf_index(want["post_probs_l"].detach(), x_ref["u_post_l"])
    assert bool(((own_l == idx_l) | (want["margin_l"] < 2e-4)).all())
    own_h = O.inverse_cdf_index(want["post_probs_h"].detach(), x_ref["u_post_h"])
    assert bool(((own_h == idx_h) | (want["margin_h"] < 2e-4)).all())
    up = mt_upstream(B, T, dims, bench_loss=False)
    want["kl_l"] = O.kl_per_sample(want["post_probs_l"], want["prior_probs_l"], True)
    want["kl_h"] = O.kl_per_sample(want["post_probs_h"], want["prior_probs_h"], True)
    want["feature"] = want["post_feature"]
    sum((got[k] * up[k]).sum() for k in up).backward()
    sum((want[k] * up[k]).sum() for k in up).backward()
    rep = H.Report(f"cfg2 MMTRSSM fp32 path B={B} T={T} vs GPU-eager fp32 oracle")
    rep.check("feature", got["feature"], want["post_feature"], rtol=1e-5, atol=2e-6)
    for k in (*MT_STATE_KEYS, "kl_l", "kl_h"):
        rep.check(k, got[k], want[k], rtol=1e-5, atol=2e-6)
    for k in MT_GRAD_IN:
        r = x_ref[k].grad
        rep.check("d " + k, x[k].grad, r, rtol=1e-4, atol=2e-5 * max(float(r.abs().max()), 1e-3))
    for k in w:
        r = w_ref[k].grad
        rep.check("d " + k, w[k].grad, r, rtol=1e-4, atol=2e-5 * max(float(r.abs().max()), 1e-3))
    rep.finish()


@pytest.mark.parametrize("precision_name", ["bf16_fused", "bf16_two_kernel"])
def test_cfg4_long_horizon_T512_drift_curve_vs_oracle(ops, precision_name):
    """BASELINE.json configs[3]: T = 512, B = 256, bf16 tensor-core gate path.  Per-step drift of every state / probability tensor
    against the fp32 oracle (teacher-forced), asserted at EVERY step against the flat-in-T bound, plus all gradients through
    512 steps of BPTT."""
    R, P = ops
    from multimodal_mtrssm_b200 import _lib

    prec = {"bf16_fused": _lib.PRECISION_BF16_FUSED, "bf16_two_kernel": _lib.PRECISION_BF16}[precision_name]
    B, T, dims = 256, 512, H.MT_DIMS
    got, w, x, want, w_ref, x_ref, inp = run_mt_both(R, P, B, T, prec, bench_loss=False)
    st = Stats(f"cfg4 MMTRSSM {precision_name} B={B} T={T} vs GPU-eager fp32 oracle (teacher-forced)")
    check_mt(st, got, w, x, want, w_ref, x_ref, inp, dims, STATE_MAX, STATE_RMS, GRAD_MAX, GRAD_RMS)
    # drift curve: max over (batch, features) per step
    f = got["feature"].detach()
    curves = {"deter_h": (f[..., :32] - want["deter_h"]).abs().amax((0, 2)), "deter_l": (f[..., 48:80] - want["deter_l"]).abs().amax((0, 2)),
              "hidden_l": (got["hidden_l"] - want["hidden_l"]).abs().amax((0, 2)),
              "post_probs_l": (got["post_probs_l"] - want["post_probs_l"]).abs().amax((0, 2, 3)),
              "post_probs_h": (got["post_probs_h"] - want["post_probs_h"]).abs().amax((0, 2, 3))}
    for k, c in curves.items():
        c = c.detach().cpu()
        st.rows[f"drift[{k}]"] = dict(t0_63=float(c[:64].max()), t64_255=float(c[64:256].max()), t256_511=float(c[256:].max()),
                                      worst_step=int(c.argmax()), ok=bool((c <= STATE_MAX).all()))
        _LOG.setdefault("cfg4 drift curves " + precision_name, {})[k] = [round(float(v), 5) for v in c]
        if not st.rows[f"drift[{k}]"]["ok"]:
            st.failed.append(f"drift[{k}]")
    st.finish()


# =====================================================================================================================
# MoPoE-MRSSM, default.yaml dims (cfg1) -- bf16 GRADIENTS of the multimodal path (forward-only before)
# =====================================================================================================================
MR_GRAD_IN = ("actions", "embed_a", "embed_v", "h0", "z0")


def run_mr_both(R, P, B, T, K, precision, D=32, gain=2.0, balancing=True):
    C = 16 // K
    params = H.make_params(H.mr_shapes(D) if D != 32 else H.MR_SHAPES, gain=gain)
    inp = H.mrssm_inputs(B, T, C, K, D=D)
    inp["u_prior"] = None
    g = torch.Generator(device="cuda").manual_seed(7)
    r = lambda *s: torch.randn(*s, generator=g, device="cuda")  # noqa: E731
    up = {"feature": r(B, T, D + 16), "kl": r(B, T), "post_probs": r(B, T, C, K), "prior_probs": r(B, T, C, K)}
    w = {k: v.cuda().requires_grad_(True) for k, v in params.items()}
    x = cuda(inp)
    for k in MR_GRAD_IN:
        x[k] = x[k].requires_grad_(True)
    got = R.mrssm_rollout(P.mrssm_weight_list(w), class_size=K, precision=precision, use_kl_balancing=balancing, **x)
    sum((got[k] * up[k]).sum() for k in up).backward()
    idx = got["feature"][..., D:].detach().reshape(B, T, C, K).argmax(-1)
    w_ref = {k: v.cuda().requires_grad_(True) for k, v in params.items()}
    x_ref = cuda(inp)
    for k in MR_GRAD_IN:
        x_ref[k] = x_ref[k].requires_grad_(True)
    want = O.mrssm_rollout(w_ref, C=C, K=K, forced_post_idx=idx, **x_ref)
    want["kl"] = O.kl_per_sample(want["post_probs"], want["prior_probs"], balancing)
    want["feature"] = want["post_feature"]
    sum((want[k] * up[k]).sum() for k in up).backward()
    torch.cuda.synchronize()
    return got, w, x, want, w_ref, x_ref, inp


def check_mr(st: Stats, got, w, x, want, w_ref, x_ref, inp, D, C, K, smax, srms, gmax, grms):
    st.state("deter", got["feature"][..., :D], want["deter"], smax, srms)
    st.state("prior_probs", got["prior_probs"], want["prior_probs"], smax, srms)
    st.state("post_probs", got["post_probs"], want["post_probs"], smax, srms)
    err = (got["kl"] - want["kl"]).abs()
    st.rows["kl"] = dict(max_abs=float(err.max()), ok=bool((err <= 5e-2 + 5e-2 * want["kl"].abs()).all()))
    if not st.rows["kl"]["ok"]:
        st.failed.append("kl")
    z = got["feature"][..., D:].detach().reshape(*got["feature"].shape[:2], C, K)
    assert bool(((z == 0) | (z == 1)).all()) and bool((z.sum(-1) == 1).all())
    u = inp["u_post"].cuda()
    assert bool(((O.inverse_cdf_index(got["post_probs"].detach(), u) == z.argmax(-1)) | (O.cdf_margin(got["post_probs"].detach(), u) < 1e-5)).all())
    for k in MR_GRAD_IN:
        st.grad("d " + k, x[k].grad, x_ref[k].grad, gmax, grms)
    for k in w:
        st.grad("d " + k.replace("rnn_to_post_projector", "post").replace("_projector", ""), w[k].grad, w_ref[k].grad, gmax, grms)


@pytest.mark.parametrize("B,T,K,balancing", [(8, 30, 4, True), (48, 8, 4, False), (4096, 30, 4, True), (33, 12, 2, True)])
def test_cfg1_mrssm_bf16_forward_and_gradients_vs_oracle(ops, B, T, K, balancing):
    """MoPoE-MRSSM default dims, bf16 tensor-core policy: states AND every gradient of the MULTIMODAL rollout (cfg1's own B = 8,
    T = 30; a bench-sized batch; ragged and K = 2 cases)."""
    R, P = ops
    from multimodal_mtrssm_b200 import _lib

    res = run_mr_both(R, P, B, T, K, _lib.PRECISION_BF16, balancing=balancing)
    st = Stats(f"cfg1 MRSSM bf16 B={B} T={T} K={K} balancing={balancing} vs GPU-eager fp32 oracle (teacher-forced)")
    check_mr(st, *res, 32, 16 // K, K, STATE_MAX, STATE_RMS, GRAD_MAX, GRAD_RMS)
    st.finish()


# =====================================================================================================================
# cfg3: MoPoE-MRSSM rollout microbench at its FULL size, B = 1024, T = 64, hidden 512
# =====================================================================================================================
def test_cfg3_wide_full_size_forward_and_gradients_vs_oracle(ops):
    """BASELINE.json configs[2] at full size against the eager fp32 oracle on the same GPU (teacher-forced): every state and
    probability, the KL, and every gradient -- not properties only."""
    R, P = ops
    from multimodal_mtrssm_b200 import _lib

    D, B, T, K = 512, 1024, 64, 4
    res = run_mr_both(R, P, B, T, K, _lib.PRECISION_BF16, D=D, gain=1.0)
    st = Stats(f"cfg3 wide MRSSM bf16 D={D} B={B} T={T} vs GPU-eager fp32 oracle (teacher-forced)")
    check_mr(st, *res, D, 16 // K, K, WIDE_STATE_MAX, WIDE_STATE_RMS, WIDE_GRAD_MAX, WIDE_GRAD_RMS)
    st.finish()
