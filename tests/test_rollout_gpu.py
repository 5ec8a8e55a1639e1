"""GPU parity: CUDA rollout kernels (through the C ABI / custom ops) vs the CPU oracle and the golden fixtures.

Tolerances (north_star): fp32 path -- 1e-5 relative (atol 1e-6 on O(1) quantities) for means/probs/latents/ELBO
terms and 1e-4 relative (atol 1e-6) for gradients accumulated over B*T; bf16 path -- stated per test below.
Discrete draws are compared exactly; uniforms are kept `eps` away from CDF boundaries of the oracle trajectory so
that rounding-level differences cannot flip a draw (helpers.*_safe_uniforms).
"""

import pytest
import torch

from oracle import rssm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

FWD_TOL = dict(rtol=1e-5, atol=2e-6)


def grad_tol(ref: torch.Tensor) -> dict:
    """Gradients are sums over up to B*T terms pushed through T steps of BPTT: compare relative to the tensor's
    scale (2e-5 * max|grad|) plus 1e-4 elementwise -- fp32 rounding of two correct implementations differs by this."""
    return dict(rtol=1e-4, atol=2e-5 * max(float(ref.abs().max()), 1e-3))


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from multimodal_mtrssm_b200 import params as P
    from multimodal_mtrssm_b200 import rollout_ops as R

    return R, P


def cuda(d):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


# ---------------------------------------------------------------------------------------------------
# MRSSM
# ---------------------------------------------------------------------------------------------------
def run_mrssm(R, P, params, inp, K, precision=0, grad=False, upstream=None, use_balancing=True):
    w = {k: v.cuda().requires_grad_(grad) for k, v in params.items()}
    x = cuda(inp)
    if grad:
        for k in ("actions", "embed_a", "embed_v", "h0", "z0"):
            x[k] = x[k].requires_grad_(True)
    out = R.mrssm_rollout(P.mrssm_weight_list(w), class_size=K, precision=precision, use_kl_balancing=use_balancing, **x)
    if grad:
        loss = sum((out[k] * upstream[k].cuda()).sum() for k in upstream)
        loss.backward()
    torch.cuda.synchronize()
    return out, w, x


def oracle_mrssm(params, inp, C, K, grad=False, upstream=None, use_balancing=True, forced=None):
    w = {k: v.clone().requires_grad_(grad) for k, v in params.items()}
    x = dict(inp)
    if grad:
        for k in ("actions", "embed_a", "embed_v", "h0", "z0"):
            x[k] = x[k].clone().requires_grad_(True)
    res = O.mrssm_rollout(w, C=C, K=K, forced_post_idx=forced, **x)
    res["kl"] = O.kl_per_sample(res["post_probs"], res["prior_probs"], use_balancing)
    res["feature"] = res["post_feature"]
    if grad:
        loss = sum((res[k] * upstream[k]).sum() for k in upstream)
        loss.backward()
    return res, w, x


def mrssm_upstream(B, T, C, K, seed=7):
    g = torch.Generator().manual_seed(seed)
    return {
        "feature": torch.randn(B, T, 48, generator=g), "kl": torch.randn(B, T, generator=g),
        "post_probs": torch.randn(B, T, C, K, generator=g), "prior_probs": torch.randn(B, T, C, K, generator=g),
        "prior_stoch": torch.randn(B, T, 16, generator=g),
    }


@pytest.mark.parametrize("B,T,K", [(37, 9, 4), (1, 1, 4), (16, 3, 2), (33, 5, 8), (8, 30, 16)])
def test_mrssm_forward_fp32(ops, B, T, K):
    R, P = ops
    C = 16 // K
    params = H.make_params(H.MR_SHAPES)
    inp = H.mrssm_inputs(B, T, C, K)
    H.mrssm_safe_uniforms(params, inp, C, K, eps=2e-4)
    want, _, _ = oracle_mrssm(params, inp, C, K)
    got, _, _ = run_mrssm(R, P, params, inp, K)
    rep = H.Report(f"mrssm fwd fp32 B={B} T={T} K={K}")
    rep.check("deter", got["feature"][..., :32], want["deter"], **FWD_TOL)
    rep.check("post_stoch", got["feature"][..., 32:], want["post_stoch"], **FWD_TOL)
    rep.check("prior_probs", got["prior_probs"], want["prior_probs"], **FWD_TOL)
    rep.check("post_probs", got["post_probs"], want["post_probs"], **FWD_TOL)
    rep.check("prior_stoch", got["prior_stoch"], want["prior_stoch"], **FWD_TOL)
    rep.check("kl", got["kl"], want["kl"], rtol=1e-5, atol=1e-6)
    rep.finish()
    assert torch.equal(got["feature"][..., 32:].cpu().reshape(B, T, C, K).argmax(-1), want["post_idx"])


@pytest.mark.parametrize("B,T,K,balancing", [(37, 9, 4, True), (5, 30, 4, False), (20, 4, 2, True)])
def test_mrssm_backward_fp32(ops, B, T, K, balancing):
    R, P = ops
    C = 16 // K
    params = H.make_params(H.MR_SHAPES)
    inp = H.mrssm_inputs(B, T, C, K)
    H.mrssm_safe_uniforms(params, inp, C, K, eps=2e-4)
    up = mrssm_upstream(B, T, C, K)
    _, w_ref, x_ref = oracle_mrssm(params, inp, C, K, grad=True, upstream=up, use_balancing=balancing)
    _, w, x = run_mrssm(R, P, params, inp, K, grad=True, upstream=up, use_balancing=balancing)
    rep = H.Report(f"mrssm bwd fp32 B={B} T={T} K={K} balancing={balancing}")
    for k in ("embed_a", "embed_v", "actions", "h0", "z0"):
        rep.check("d " + k, x[k].grad, x_ref[k].grad, **grad_tol(x_ref[k].grad))
    for k in w:
        rep.check("d " + k.replace("rnn_to_", "").replace("_projector", ""), w[k].grad, w_ref[k].grad, **grad_tol(w_ref[k].grad))
    rep.finish()


def test_mrssm_only_kl_loss_and_no_prior_sample(ops):
    """Upstream gradient only through kl (d_feature absent) and without the prior's own draw."""
    R, P = ops
    B, T, C, K = 19, 6, 4, 4
    params = H.make_params(H.MR_SHAPES)
    inp = H.mrssm_inputs(B, T, C, K)
    H.mrssm_safe_uniforms(params, inp, C, K, eps=2e-4)
    inp["u_prior"] = None
    up = {"kl": torch.full((B, T), 1.0 / (B * T))}
    _, w_ref, _ = oracle_mrssm(params, inp, C, K, grad=True, upstream=up)
    out, w, _ = run_mrssm(R, P, params, inp, K, grad=True, upstream=up)
    assert out["prior_stoch"] is None
    rep = H.Report("mrssm kl-only backward")
    for k in w:
        rep.check("d " + k, w[k].grad, w_ref[k].grad, **grad_tol(w_ref[k].grad))
    rep.finish()


def test_mrssm_imagine_fp32(ops):
    R, P = ops
    B, T, C, K = 21, 11, 4, 4
    params = H.make_params(H.MR_SHAPES)
    inp = H.mrssm_inputs(B, T, C, K)
    g = torch.Generator().manual_seed(5)
    u = torch.rand(B, T, C, generator=g)
    for _ in range(50):
        want = O.mrssm_imagine(params, actions=inp["actions"], h0=inp["h0"], z0=inp["z0"], u=u, C=C, K=K)
        bad = O.cdf_margin(want["probs"], u) < 2e-4
        if not bad.any():
            break
        u[bad] = torch.rand(int(bad.sum()), generator=g)
    w = {k: v.cuda() for k, v in params.items()}
    got = R.mrssm_imagine(P.mrssm_weight_list(w), actions=inp["actions"].cuda(), h0=inp["h0"].cuda(), z0=inp["z0"].cuda(),
                          u=u.cuda(), class_size=K)
    rep = H.Report("mrssm imagine fp32")
    rep.check("deter", got["feature"][..., :32], want["deter"], **FWD_TOL)
    rep.check("stoch", got["feature"][..., 32:], want["stoch"], **FWD_TOL)
    rep.check("probs", got["probs"], want["probs"], **FWD_TOL)
    rep.finish()


@pytest.mark.parametrize("fixture", H.MRSSM_GOLDEN)
def test_mrssm_golden_fixture(ops, golden_dir, fixture):
    """The CUDA path against vectors produced by the reference's own code (tests/golden/make_golden.py): B = 5, T = 7 and
    default.yaml's own B = 8, T = 30 (BASELINE.json configs[0])."""
    R, P = ops
    g = torch.load(golden_dir / fixture)
    dims, inp, out = g["dims"], g["inputs"], g["outputs"]
    w = {k: v.cuda().requires_grad_(True) for k, v in g["params"].items() if not k.startswith("representation.")}
    x = {k: inp[k].cuda().requires_grad_(True) for k in ("embed_a", "embed_v", "h0", "z0")}
    res = R.mrssm_rollout(P.mrssm_weight_list(w), actions=inp["actions"].cuda(), u_post=inp["u_post"].cuda(),
                          u_prior=inp["u_prior"].cuda(), class_size=dims["K"], **x)
    rep = H.Report("mrssm golden (reference code) vs CUDA")
    rep.check("post_feature", res["feature"], out["post_feature"], **FWD_TOL)
    rep.check("post_probs", res["post_probs"], out["post_probs"], **FWD_TOL)
    rep.check("prior_probs", res["prior_probs"], out["prior_probs"], **FWD_TOL)
    rep.check("prior_stoch", res["prior_stoch"], out["prior_stoch"], **FWD_TOL)
    kl = res["kl"].mean() * dims["kl_coeff"]
    rep.check("kl", kl, g["loss"]["kl"], rtol=1e-5, atol=1e-7)
    ((res["feature"] * g["upstream"]["d_post_feature"].cuda()).sum() + kl).backward()
    rep.check("d embed_a", x["embed_a"].grad, g["grads"]["embed_a"], **grad_tol(g["grads"]["embed_a"]))
    rep.check("d embed_v", x["embed_v"].grad, g["grads"]["embed_v"], **grad_tol(g["grads"]["embed_v"]))
    rep.check("d z0", x["z0"].grad, g["grads"]["z0"], **grad_tol(g["grads"]["z0"]))
    # golden h0 / prior-projector grads include the initial_state path through z0 (core.py:133-135); compare the rest
    for k, ref in g["grads"]["params"].items():
        if k.startswith(("representation.", "transition.rnn_to_prior_projector")):
            continue
        rep.check("d " + k, w[k].grad, ref, **grad_tol(ref))
    rep.finish()


def test_mrssm_bf16_teacher_forced(ops):
    """bf16 tensor-core path.  Stated tolerance: 3e-2 absolute on O(1) states/probs (bf16 operands, 8-bit mantissa,
    error compounding over T steps of recurrence), samples self-consistent with the kernel's own probabilities."""
    R, P = ops
    B, T, C, K = 64, 16, 4, 4
    params = H.make_params(H.MR_SHAPES)
    inp = H.mrssm_inputs(B, T, C, K)
    got, _, _ = run_mrssm(R, P, params, inp, K, precision=1)
    idx = got["feature"][..., 32:].cpu().reshape(B, T, C, K).argmax(-1)
    want, _, _ = oracle_mrssm(params, inp, C, K, forced=idx)
    rep = H.Report("mrssm fwd bf16 (teacher-forced on the kernel's draws)")
    rep.check("deter", got["feature"][..., :32], want["deter"], rtol=0, atol=3e-2)
    rep.check("prior_probs", got["prior_probs"], want["prior_probs"], rtol=0, atol=3e-2)
    rep.check("post_probs", got["post_probs"], want["post_probs"], rtol=0, atol=3e-2)
    rep.finish()
    # the kernel's draw is the inverse-CDF draw of ITS OWN probabilities
    self_idx = O.inverse_cdf_index(got["post_probs"].cpu(), inp["u_post"])
    margin = O.cdf_margin(got["post_probs"].cpu(), inp["u_post"])
    assert bool(((self_idx == idx) | (margin < 1e-5)).all())
    onehot = got["feature"][..., 32:].cpu().reshape(B, T, C, K)
    assert bool(((onehot == 0) | (onehot == 1)).all()) and bool((onehot.sum(-1) == 1).all())


# ---------------------------------------------------------------------------------------------------
# unimodal rollout (BaseRSSM.rollout_representation, core.py:137-168): dims.unimodal = 1
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T,K,precision", [(37, 9, 4, 0), (8, 30, 2, 0), (20, 6, 4, 1)])
def test_mrssm_unimodal_matches_oracle(ops, B, T, K, precision):
    """One Representation.forward per step, no fusion: forward and every gradient against the oracle's unimodal switch.  The
    vision slot (weights and embedding) must not influence the result and must get exactly zero gradients.  fp32 path: the
    module tolerances; bf16 path: teacher-forced on the kernel's draws, 3e-2 / 2e-2 of scale as for the multimodal kernels."""
    R, P = ops
    C = 16 // K
    params = H.make_params(H.MR_SHAPES)
    inp = H.mrssm_inputs(B, T, C, K)
    H.mrssm_safe_uniforms(params, inp, C, K, eps=2e-4, unimodal=True)
    up = mrssm_upstream(B, T, C, K)
    w = {k: v.cuda().requires_grad_(True) for k, v in params.items()}
    x = cuda(inp)
    for k in ("actions", "embed_a", "embed_v", "h0", "z0"):
        x[k] = x[k].requires_grad_(True)
    got = R.mrssm_rollout(P.mrssm_weight_list(w), class_size=K, precision=precision, unimodal=True, **x)
    sum((got[k] * up[k].cuda()).sum() for k in up).backward()
    forced = None if precision == 0 else got["feature"][..., 32:].detach().cpu().reshape(B, T, C, K).argmax(-1)
    w_ref = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    x_ref = {k: (v.clone().requires_grad_(True) if k in ("actions", "embed_a", "embed_v", "h0", "z0") else v) for k, v in inp.items()}
    want = O.mrssm_rollout(w_ref, C=C, K=K, forced_post_idx=forced, unimodal=True, **x_ref)
    want["kl"] = O.kl_per_sample(want["post_probs"], want["prior_probs"], True)
    want["feature"] = want["post_feature"]
    if precision != 0:  # the prior's own draw is not teacher-forced: leave it out of the bf16 comparison
        up = {k: v for k, v in up.items() if k != "prior_stoch"}
        w2 = {k: v.cuda().requires_grad_(True) for k, v in params.items()}
        x2 = cuda(inp)
        for k in ("actions", "embed_a", "embed_v", "h0", "z0"):
            x2[k] = x2[k].requires_grad_(True)
        got = R.mrssm_rollout(P.mrssm_weight_list(w2), class_size=K, precision=precision, unimodal=True, **x2)
        sum((got[k] * up[k].cuda()).sum() for k in up).backward()
        w, x = w2, x2
    sum((want[k] * up[k]).sum() for k in up).backward()
    rep = H.Report(f"mrssm unimodal B={B} T={T} K={K} precision={precision}")
    ftol = FWD_TOL if precision == 0 else dict(rtol=0, atol=3e-2)
    rep.check("deter", got["feature"][..., :32], want["deter"], **ftol)
    rep.check("post_probs", got["post_probs"], want["post_probs"], **ftol)
    rep.check("prior_probs", got["prior_probs"], want["prior_probs"], **ftol)
    rep.check("kl", got["kl"], want["kl"], **(dict(rtol=1e-5, atol=1e-6) if precision == 0 else dict(rtol=0, atol=3e-2)))
    if precision == 0:
        assert torch.equal(got["feature"][..., 32:].cpu().reshape(B, T, C, K).argmax(-1), want["post_idx"])
    gt = grad_tol if precision == 0 else (lambda r: dict(rtol=0, atol=2e-2 * max(float(r.abs().max()), 1e-3)))
    for k in ("embed_a", "actions", "h0", "z0"):
        rep.check("d " + k, x[k].grad, x_ref[k].grad, **gt(x_ref[k].grad))
    for k in w:
        if k.startswith("vision_representation"):
            continue
        rep.check("d " + k, w[k].grad, w_ref[k].grad, **gt(w_ref[k].grad))
    rep.finish()
    assert float(x["embed_v"].grad.abs().max()) == 0.0
    assert all(float(w[k].grad.abs().max()) == 0.0 for k in w if k.startswith("vision_representation"))


def test_unimodal_golden_fixture(ops, golden_dir):
    """SURVEY §8 row a6 against the reference's OWN BaseRSSM.rollout_representation / rollout_transition (models/core.py:137-185,
    run through a concrete subclass by make_golden.golden_unimodal): states, KL, every gradient, imagination."""
    R, P = ops
    g = torch.load(golden_dir / "rssm_unimodal.pt")
    dims, inp, out = g["dims"], g["inputs"], g["outputs"]
    ren = lambda k: k.replace("representation.", "audio_representation.", 1) if k.startswith("representation.") else k  # noqa: E731
    w = {ren(k): v.cuda().requires_grad_(True) for k, v in g["params"].items()}
    for k in list(w):  # the unused vision slot
        if k.startswith("audio_representation."):
            w[k.replace("audio_", "vision_", 1)] = w[k].detach().clone().requires_grad_(True)
    x = {"embed_a": inp["embed"].cuda().requires_grad_(True), "h0": inp["h0"].cuda().requires_grad_(True),
         "z0": inp["z0"].cuda().requires_grad_(True)}
    res = R.mrssm_rollout(P.mrssm_weight_list(w), actions=inp["actions"].cuda(), embed_v=inp["embed"].cuda(), u_post=inp["u_post"].cuda(),
                          u_prior=inp["u_prior"].cuda(), class_size=dims["K"], unimodal=True, **x)
    rep = H.Report("unimodal golden (reference BaseRSSM) vs CUDA")
    rep.check("post_feature", res["feature"], out["post_feature"], **FWD_TOL)
    rep.check("post_probs", res["post_probs"], out["post_probs"], **FWD_TOL)
    rep.check("prior_probs", res["prior_probs"], out["prior_probs"], **FWD_TOL)
    rep.check("prior_stoch", res["prior_stoch"], out["prior_stoch"], **FWD_TOL)
    kl = res["kl"].mean() * dims["kl_coeff"]
    rep.check("kl", kl, g["loss"]["kl"], rtol=1e-5, atol=1e-7)
    ((res["feature"] * g["upstream"]["d_post_feature"].cuda()).sum() + kl).backward()
    rep.check("d embed", x["embed_a"].grad, g["grads"]["embed"], **grad_tol(g["grads"]["embed"]))
    rep.check("d z0", x["z0"].grad, g["grads"]["z0"], **grad_tol(g["grads"]["z0"]))
    for k, ref in g["grads"]["params"].items():
        if k.startswith("transition.rnn_to_prior_projector"):  # golden adds the initial_state path through z0 (core.py:133-135)
            continue
        rep.check("d " + k, w[ren(k)].grad, ref, **grad_tol(ref))
    im = g["imagine"]
    got = R.mrssm_imagine(P.mrssm_weight_list({k: v.detach() for k, v in w.items()}), actions=im["actions"].cuda(),
                          h0=out["deter"][:, -1].cuda(), z0=out["post_stoch"][:, -1].round().cuda(), u=im["u"].cuda(), class_size=dims["K"])
    rep.check("imagine deter", got["feature"][..., :32], im["deter"], **FWD_TOL)
    rep.check("imagine stoch", got["feature"][..., 32:], im["stoch"], **FWD_TOL)
    rep.check("imagine probs", got["probs"], im["probs"], **FWD_TOL)
    rep.finish()


# ---------------------------------------------------------------------------------------------------
# MMTRSSM
# ---------------------------------------------------------------------------------------------------
MT_GRAD_IN = ("actions", "embed_a", "embed_v", "deter_h0", "deter_l0", "hidden_h0", "hidden_l0", "stoch_h0", "stoch_l0")


def run_mtrssm(R, P, params, inp, dims, precision=0, grad=False, upstream=None, use_balancing=True):
    w = {k: v.cuda().requires_grad_(grad) for k, v in params.items()}
    x = cuda(inp)
    if grad:
        for k in MT_GRAD_IN:
            x[k] = x[k].requires_grad_(True)
    out = R.mtrssm_rollout(P.mtrssm_weight_list(w), class_size_l=dims["KL"], class_size_h=dims["KH"], l_tau=dims["l_tau"],
                           h_tau=dims["h_tau"], precision=precision, use_kl_balancing=use_balancing, **x)
    if grad:
        sum((out[k] * upstream[k].cuda()).sum() for k in upstream).backward()
    torch.cuda.synchronize()
    return out, w, x


def oracle_mtrssm(params, inp, dims, grad=False, upstream=None, use_balancing=True, forced=(None, None)):
    w = {k: v.clone().requires_grad_(grad) for k, v in params.items()}
    x = dict(inp)
    if grad:
        for k in MT_GRAD_IN:
            x[k] = x[k].clone().requires_grad_(True)
    res = O.mtrssm_rollout(w, dims=dims, forced_idx_l=forced[0], forced_idx_h=forced[1], **x)
    res["kl_l"] = O.kl_per_sample(res["post_probs_l"], res["prior_probs_l"], use_balancing)
    res["kl_h"] = O.kl_per_sample(res["post_probs_h"], res["prior_probs_h"], use_balancing)
    res["feature"] = res["post_feature"]
    if grad:
        sum((res[k] * upstream[k]).sum() for k in upstream).backward()
    return res, w, x


def mtrssm_upstream(B, T, dims, seed=7):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    return {
        "feature": r(B, T, 96), "kl_l": r(B, T), "kl_h": r(B, T),
        "post_probs_l": r(B, T, dims["CL"], dims["KL"]), "post_probs_h": r(B, T, dims["CH"], dims["KH"]),
        "prior_probs_l": r(B, T, dims["CL"], dims["KL"]), "prior_probs_h": r(B, T, dims["CH"], dims["KH"]),
        "prior_stoch_l": r(B, T, 16), "prior_stoch_h": r(B, T, 16),
        # gradients into the MTRNN.hidden outputs (the reference's autograd carries them, mmtrssm/mopoe_mmtrssm/core.py:472-473)
        "hidden_h": r(B, T, 32), "hidden_l": r(B, T, 32),
    }


MT_FWD_KEYS = ("hidden_h", "hidden_l", "prior_probs_h", "prior_probs_l", "post_probs_h", "post_probs_l", "prior_stoch_h",
               "prior_stoch_l", "kl_l", "kl_h")


@pytest.mark.parametrize("B,T,dims", [
    (37, 9, H.MT_DIMS), (1, 1, H.MT_DIMS), (16, 30, H.MT_DIMS),
    (18, 5, dict(CL=4, KL=4, CH=4, KH=4, l_tau=1.5, h_tau=8.0)),
    (9, 4, dict(CL=8, KL=2, CH=8, KH=2, l_tau=2.0, h_tau=3.0)),
])
def test_mtrssm_forward_fp32(ops, B, T, dims):
    R, P = ops
    params = H.make_params(H.MT_SHAPES)
    inp = H.mtrssm_inputs(B, T, dims)
    H.mtrssm_safe_uniforms(params, inp, dims, eps=2e-4)
    want, _, _ = oracle_mtrssm(params, inp, dims)
    got, _, _ = run_mtrssm(R, P, params, inp, dims)
    rep = H.Report(f"mtrssm fwd fp32 B={B} T={T} {dims}")
    rep.check("feature", got["feature"], want["post_feature"], **FWD_TOL)
    for k in MT_FWD_KEYS:
        rep.check(k, got[k], want[k], **FWD_TOL)
    rep.finish()


@pytest.mark.parametrize("B,T,balancing", [(37, 9, True), (6, 30, False)])
def test_mtrssm_backward_fp32(ops, B, T, balancing):
    R, P = ops
    dims = H.MT_DIMS
    params = H.make_params(H.MT_SHAPES)
    inp = H.mtrssm_inputs(B, T, dims)
    H.mtrssm_safe_uniforms(params, inp, dims, eps=2e-4)
    up = mtrssm_upstream(B, T, dims)
    _, w_ref, x_ref = oracle_mtrssm(params, inp, dims, grad=True, upstream=up, use_balancing=balancing)
    _, w, x = run_mtrssm(R, P, params, inp, dims, grad=True, upstream=up, use_balancing=balancing)
    rep = H.Report(f"mtrssm bwd fp32 B={B} T={T} balancing={balancing}")
    for k in MT_GRAD_IN:
        rep.check("d " + k, x[k].grad, x_ref[k].grad, **grad_tol(x_ref[k].grad))
    for k in w:
        rep.check("d " + k.replace("rnn_to_post_projector", "post"), w[k].grad, w_ref[k].grad, **grad_tol(w_ref[k].grad))
    rep.finish()


def test_mtrssm_imagine_fp32(ops):
    R, P = ops
    B, T, dims = 21, 11, H.MT_DIMS
    params = H.make_params(H.MT_SHAPES)
    inp = H.mtrssm_inputs(B, T, dims)
    state = {k: inp[k] for k in ("deter_h0", "deter_l0", "hidden_h0", "hidden_l0", "stoch_h0", "stoch_l0")}
    g = torch.Generator().manual_seed(5)
    u_l, u_h = torch.rand(B, T, dims["CL"], generator=g), torch.rand(B, T, dims["CH"], generator=g)
    for _ in range(50):
        want = O.mtrssm_imagine(params, actions=inp["actions"], u_l=u_l, u_h=u_h, dims=dims, **state)
        bad_l, bad_h = O.cdf_margin(want["probs_l"], u_l) < 2e-4, O.cdf_margin(want["probs_h"], u_h) < 2e-4
        if not (bad_l.any() or bad_h.any()):
            break
        u_l[bad_l] = torch.rand(int(bad_l.sum()), generator=g)
        u_h[bad_h] = torch.rand(int(bad_h.sum()), generator=g)
    w = {k: v.cuda() for k, v in params.items()}
    got = R.mtrssm_imagine(P.mtrssm_weight_list(w), actions=inp["actions"].cuda(), u_l=u_l.cuda(), u_h=u_h.cuda(),
                           class_size_l=dims["KL"], class_size_h=dims["KH"], l_tau=dims["l_tau"], h_tau=dims["h_tau"], **cuda(state))
    rep = H.Report("mtrssm imagine fp32")
    f = got["feature"]
    for name, a, b in (("deter_h", f[..., :32], want["deter_h"]), ("stoch_h", f[..., 32:48], want["stoch_h"]),
                       ("deter_l", f[..., 48:80], want["deter_l"]), ("stoch_l", f[..., 80:], want["stoch_l"]),
                       ("hidden_h", got["hidden_h"], want["hidden_h"]), ("hidden_l", got["hidden_l"], want["hidden_l"]),
                       ("probs_h", got["probs_h"], want["probs_h"]), ("probs_l", got["probs_l"], want["probs_l"])):
        rep.check(name, a, b, **FWD_TOL)
    rep.finish()


@pytest.mark.parametrize("fixture", H.MTRSSM_GOLDEN)
def test_mtrssm_golden_fixture(ops, golden_dir, fixture):
    """B = 5, T = 7 and default.yaml's own B = 8, T = 30 (BASELINE.json configs[1])."""
    R, P = ops
    g = torch.load(golden_dir / fixture)
    dims, inp, out = g["dims"], g["inputs"], g["outputs"]
    w = {k: v.cuda().requires_grad_(True) for k, v in g["params"].items()}
    names = ("embed_a", "embed_v", "deter_h0", "deter_l0", "hidden_h0", "hidden_l0", "stoch_h0", "stoch_l0")
    x = {k: inp[k].cuda().requires_grad_(True) for k in names}
    u = {k: inp[k].cuda() for k in ("u_post_l", "u_post_h", "u_prior_l", "u_prior_h")}
    res = R.mtrssm_rollout(P.mtrssm_weight_list(w), actions=inp["actions"].cuda(), class_size_l=dims["KL"], class_size_h=dims["KH"],
                           l_tau=dims["l_tau"], h_tau=dims["h_tau"], **x, **u)
    rep = H.Report("mtrssm golden (reference code) vs CUDA")
    rep.check("post_feature", res["feature"], out["post_feature"], **FWD_TOL)
    for k in ("hidden_h", "hidden_l", "post_probs_h", "post_probs_l", "prior_probs_h", "prior_probs_l", "prior_stoch_h", "prior_stoch_l"):
        rep.check(k, res[k], out[k], **FWD_TOL)
    kl_l = res["kl_l"].mean() * dims["kl_coeff"]
    kl_h = res["kl_h"].mean() * dims["kl_coeff"] * dims["w_kl_h"]
    rep.check("kl", kl_l, g["loss"]["kl"], rtol=1e-5, atol=1e-7)
    rep.check("kl_h", kl_h, g["loss"]["kl_h"], rtol=1e-5, atol=1e-7)
    ((res["feature"] * g["upstream"]["d_post_feature"].cuda()).sum() + kl_l + kl_h).backward()
    for k in ("embed_a", "embed_v", "stoch_h0", "stoch_l0"):
        rep.check("d " + k, x[k].grad, g["grads"][k], **grad_tol(g["grads"][k]))
    for k, ref in g["grads"]["params"].items():
        if k.startswith(("l_prior", "h_prior")):  # golden adds the initial_state path through stoch_*0
            continue
        rep.check("d " + k, w[k].grad, ref, **grad_tol(ref))
    rep.finish()


def test_mtrssm_bf16_teacher_forced(ops):
    """bf16 tensor-core path; stated tolerance 3e-2 absolute on O(1) states/probs (see the MRSSM twin)."""
    R, P = ops
    B, T, dims = 64, 16, H.MT_DIMS
    params = H.make_params(H.MT_SHAPES)
    inp = H.mtrssm_inputs(B, T, dims)
    got, _, _ = run_mtrssm(R, P, params, inp, dims, precision=1)
    f = got["feature"].cpu()
    idx_h = f[..., 32:48].reshape(B, T, dims["CH"], dims["KH"]).argmax(-1)
    idx_l = f[..., 80:].reshape(B, T, dims["CL"], dims["KL"]).argmax(-1)
    want, _, _ = oracle_mtrssm(params, inp, dims, forced=(idx_l, idx_h))
    rep = H.Report("mtrssm fwd bf16 (teacher-forced)")
    rep.check("deter_h", f[..., :32], want["deter_h"], rtol=0, atol=3e-2)
    rep.check("deter_l", f[..., 48:80], want["deter_l"], rtol=0, atol=3e-2)
    for k in ("hidden_h", "hidden_l", "prior_probs_h", "prior_probs_l", "post_probs_h", "post_probs_l"):
        rep.check(k, got[k], want[k], rtol=0, atol=3e-2)
    rep.finish()


@pytest.mark.parametrize("B,T,dims", [
    (37, 9, H.MT_DIMS), (200, 12, H.MT_DIMS), (16, 1, H.MT_DIMS), (1, 3, H.MT_DIMS), (20, 131, H.MT_DIMS),
    (70, 5, dict(CL=4, KL=4, CH=4, KH=4, l_tau=1.5, h_tau=8.0)),
    (33, 4, dict(CL=8, KL=2, CH=8, KH=2, l_tau=2.0, h_tau=3.0)),
])
def test_mtrssm_bf16_fused_backward_matches_two_kernel_backward(ops, B, T, dims):
    """RSSM_PRECISION_BF16_FUSED (BPTT + weight gradients in one kernel, tcgen05 / TMEM accumulators) against
    RSSM_PRECISION_BF16 (BPTT kernel + mma.sync weight-gradient kernel): the forward is bit-identical, the data gradients run
    the same arithmetic up to the fp32 summation order and the bf16 operand roundings that order can flip (6e-3 of scale, data
    gradients -- measured 4.4e-3 at T = 131 with every output incl. the hiddens receiving a gradient; 2e-3 of scale, weight gradients)."""
    R, P = ops
    params = H.make_params(H.MT_SHAPES)
    inp = H.mtrssm_inputs(B, T, dims)
    up = mtrssm_upstream(B, T, dims)
    o1, w1, x1 = run_mtrssm(R, P, params, inp, dims, precision=1, grad=True, upstream=up)
    o2, w2, x2 = run_mtrssm(R, P, params, inp, dims, precision=2, grad=True, upstream=up)
    for k in ("feature", *MT_FWD_KEYS):
        assert torch.equal(o1[k], o2[k]), k
    rep = H.Report(f"mtrssm bf16 fused vs two-kernel backward B={B} T={T} {dims}")
    # same arithmetic, but the two-warp fused kernel sums the d deter_l contributions in another order: an fp32 ulp can flip
    # the bf16 rounding (2^-9) of an MMA operand downstream, so the paths agree to a few bf16 ulps, not bit for bit
    for k in MT_GRAD_IN:
        scale = float(x1[k].grad.abs().max())
        rep.check("d " + k, x2[k].grad, x1[k].grad, rtol=0, atol=6e-3 * max(scale, 1e-3))
    for k in w1:
        scale = float(w1[k].grad.abs().max())
        rep.check("d " + k.replace("rnn_to_post_projector", "post"), w2[k].grad, w1[k].grad, rtol=0, atol=2e-3 * max(scale, 1e-3))
    rep.finish()


@pytest.mark.parametrize("B,T,prior", [(45, 7, True), (16, 3, False)])
def test_mtrssm_grouped_output_rows_match_dense_outputs(ops, B, T, prior):
    """The bf16 fused policy through `mtrssm_rollout` writes the outputs of a (b,t) into one 1 KB row (RssmMtrssmOutputs.ld_* = 256,
    strided views handed out); the same policy through the raw op keeps one dense tensor per output (ld_* = 0).  Same kernels, only
    the row pitches differ: forward bit-identical; gradients equal up to the atomic summation order of the weight gradients."""
    R, P = ops
    dims = H.MT_DIMS
    params = H.make_params(H.MT_SHAPES)
    inp = H.mtrssm_inputs(B, T, dims)
    if not prior:
        inp = {k: v for k, v in inp.items() if not k.startswith("u_prior")}
    up = mtrssm_upstream(B, T, dims)
    if not prior:
        up = {k: v for k, v in up.items() if not k.startswith("prior_stoch")}
    og, wg, xg = run_mtrssm(R, P, params, inp, dims, precision=2, grad=True, upstream=up)
    assert og["feature"].stride(-2) == 256 and og["post_probs_l"].data_ptr() == og["feature"].data_ptr() + 4 * 208
    assert (og["prior_stoch_h"] is None) == (not prior)
    # dense: the autograd-registered custom op with natural pitches
    w = {k: v.cuda().requires_grad_(True) for k, v in params.items()}
    x = cuda(inp)
    for k in MT_GRAD_IN:
        x[k] = x[k].requires_grad_(True)
    st = [x[k] for k in ("deter_h0", "deter_l0", "hidden_h0", "hidden_l0", "stoch_h0", "stoch_l0")]
    res = R.mtrssm_rollout_op(P.mtrssm_weight_list(w), x["actions"], x["embed_a"], x["embed_v"], st, x["u_post_l"], x["u_post_h"],
                              x.get("u_prior_l"), x.get("u_prior_h"), dims["KL"], dims["KH"], dims["l_tau"], dims["h_tau"], 2, 0.2, 0.8,
                              True, False)
    names = ("feature", "hidden_h", "hidden_l", "prior_probs_h", "prior_probs_l", "post_probs_h", "post_probs_l", "prior_stoch_h",
             "prior_stoch_l", "kl_l", "kl_h")
    od = dict(zip(names, res[:-1]))
    assert od["feature"].is_contiguous()
    sum((od[k] * up[k].cuda()).sum() for k in up).backward()
    torch.cuda.synchronize()
    for k in names:
        if og[k] is not None:
            assert torch.equal(og[k], od[k]), k
    rep = H.Report(f"mtrssm grouped rows vs dense outputs B={B} T={T} prior={prior}")
    for k in MT_GRAD_IN:
        assert torch.equal(xg[k].grad, x[k].grad), k
    for k in w:
        scale = float(w[k].grad.abs().max())
        rep.check("d " + k, wg[k].grad, w[k].grad, rtol=0, atol=1e-5 * max(scale, 1e-3))
    rep.finish()


def test_mtrssm_fwd2_matches_one_warp_kernel(ops, monkeypatch):
    """The two-warps-per-tile forward (mtrssm_fwd2.cu, the bf16 default) against the one-warp kernel (RSSM_FWD_ONE_WARP=1): same
    arithmetic and operand roundings, one fp32 summation order differs, so single steps agree to fp32 rounding of the bf16-operand
    dot products (2e-6 absolute on O(1) values) and the draws coincide away from CDF knife edges; multi-step rollouts are covered
    against the oracle (teacher-forced) by the tests above and in test_bench_configs_gpu.py.  Ragged batch, prior draws on."""
    R, P = ops
    B, T, dims = 333, 1, H.MT_DIMS
    params = H.make_params(H.MT_SHAPES)
    inp = H.mtrssm_inputs(B, T, dims)
    two, _, _ = run_mtrssm(R, P, params, inp, dims, precision=1)
    monkeypatch.setenv("RSSM_FWD_ONE_WARP", "1")
    one, _, _ = run_mtrssm(R, P, params, inp, dims, precision=1)
    monkeypatch.delenv("RSSM_FWD_ONE_WARP")
    rep = H.Report("mtrssm fwd2 (two warps per tile) vs one-warp kernel, one step")
    for k in ("hidden_h", "hidden_l", "prior_probs_h", "prior_probs_l", "post_probs_h", "kl_l", "kl_h"):
        rep.check(k, two[k], one[k], rtol=1e-5, atol=2e-6)
    # the modality heads' first layer sums its two halves in the other order: an fp32 ulp there can flip the bf16 rounding of one
    # hidden unit (2^-9 relative) in front of the second layer -> a few 1e-6 on the fused posterior (measured 7.8e-6)
    rep.check("post_probs_l", two["post_probs_l"], one["post_probs_l"], rtol=0, atol=5e-5)
    rep.check("deter_h", two["feature"][..., :32], one["feature"][..., :32], rtol=1e-5, atol=2e-6)
    rep.check("deter_l", two["feature"][..., 48:80], one["feature"][..., 48:80], rtol=1e-5, atol=2e-6)
    rep.finish()
    for key, probs, u in (("feature", "post_probs_l", "u_post_l"), ("prior_stoch_l", "prior_probs_l", "u_prior_l"),
                          ("prior_stoch_h", "prior_probs_h", "u_prior_h")):
        a = two[key][..., 80:] if key == "feature" else two[key]
        b = one[key][..., 80:] if key == "feature" else one[key]
        same = (a == b).reshape(B, T, -1).all(-1).cpu()
        margin = O.cdf_margin(one[probs].cpu(), inp[u]).amin(-1)
        assert bool((same | (margin < 1e-5)).all()), key


def test_mtrssm_fused_backward_two_warp_variant_matches_three_warp_kernel(ops, monkeypatch):
    """RSSM_BWD_TWO_WARP=1 selects the two-warps-per-tile instantiation of the fused backward (kept for A/B measurements): same
    arithmetic per row, so data gradients agree to fp32 rounding and weight gradients to the summation order of the TMEM
    accumulation (1e-5 of each tensor's scale).  Ragged batch, several tile groups, prior draws on."""
    R, P = ops
    B, T, dims = 333, 7, H.MT_DIMS
    params = H.make_params(H.MT_SHAPES)
    inp = H.mtrssm_inputs(B, T, dims)
    up = mtrssm_upstream(B, T, dims)
    _, w3, x3 = run_mtrssm(R, P, params, inp, dims, precision=2, grad=True, upstream=up)
    monkeypatch.setenv("RSSM_BWD_TWO_WARP", "1")
    _, w2, x2 = run_mtrssm(R, P, params, inp, dims, precision=2, grad=True, upstream=up)
    monkeypatch.delenv("RSSM_BWD_TWO_WARP")
    rep = H.Report("mtrssm fused backward: two-warp variant vs three-warp kernel")
    for k in MT_GRAD_IN:
        rep.check("d " + k, x2[k].grad, x3[k].grad, rtol=0, atol=1e-5 * max(float(x3[k].grad.abs().max()), 1e-3))
    for k in w3:
        if w3[k].grad is None:
            assert w2[k].grad is None, k
            continue
        rep.check("d " + k, w2[k].grad, w3[k].grad, rtol=0, atol=1e-5 * max(float(w3[k].grad.abs().max()), 1e-3))
    rep.finish()


@pytest.mark.parametrize("precision", [1, 2])
def test_mtrssm_bf16_backward_vs_oracle(ops, precision):
    """Gradients of the bf16 tensor-core paths against the fp32 oracle, teacher-forced on the kernel's own draws (as in
    test_mtrssm_bf16_teacher_forced).  Stated bf16 tolerance for gradients: 2e-2 of each tensor's scale."""
    R, P = ops
    B, T, dims = 48, 8, H.MT_DIMS
    params = H.make_params(H.MT_SHAPES)
    inp = H.mtrssm_inputs(B, T, dims)
    inp["u_prior_l"] = inp["u_prior_h"] = None  # the prior samples are not part of the training loss
    up = {k: v for k, v in mtrssm_upstream(B, T, dims).items() if not k.startswith("prior_stoch")}
    got, w, x = run_mtrssm(R, P, params, inp, dims, precision=precision, grad=True, upstream=up)
    f = got["feature"].detach().cpu()
    idx_h = f[..., 32:48].reshape(B, T, dims["CH"], dims["KH"]).argmax(-1)
    idx_l = f[..., 80:].reshape(B, T, dims["CL"], dims["KL"]).argmax(-1)
    _, w_ref, x_ref = oracle_mtrssm(params, inp, dims, grad=True, upstream=up, forced=(idx_l, idx_h))
    rep = H.Report(f"mtrssm bwd bf16 (precision {precision}) vs teacher-forced oracle")
    for k in MT_GRAD_IN:
        rep.check("d " + k, x[k].grad, x_ref[k].grad, rtol=0, atol=2e-2 * float(x_ref[k].grad.abs().max()))
    for k in w:
        rep.check("d " + k.replace("rnn_to_post_projector", "post"), w[k].grad, w_ref[k].grad, rtol=0,
                  atol=2e-2 * float(w_ref[k].grad.abs().max()))
    rep.finish()


@pytest.mark.parametrize("precision", [1, 2])
def test_mtrssm_bf16_extreme_logits_take_the_log_domain_fallback(ops, precision):
    """The bf16 policies evaluate the MoPoE fusion in the probability domain (frag.cuh: q = s / sum_group(s), s = pa + pv + pa pv)
    and fall back to the log-domain formulas when a whole group underflows.  With the modality heads' output layers scaled by
    1000 the flat log-probabilities span hundreds of nats in BOTH modalities, so the fallback runs on most rows: everything must
    stay finite and normalised, the draws one-hot, the gradients finite, and wherever the fp32 oracle (teacher-forced) is certain
    about a group (p > 0.999) the kernel must pick the same class."""
    R, P = ops
    B, T, dims = 80, 6, H.MT_DIMS
    params = H.make_params(H.MT_SHAPES)
    big = [k for k in params if ("audio" in k or "vision" in k) and k.endswith("2.weight")]
    assert len(big) == 2, list(params)
    for k in big:
        params[k] = params[k] * 1000.0
    inp = H.mtrssm_inputs(B, T, dims)
    up = {k: v for k, v in mtrssm_upstream(B, T, dims).items()}
    got, w, x = run_mtrssm(R, P, params, inp, dims, precision=precision, grad=True, upstream=up)
    for k, v in got.items():
        assert bool(torch.isfinite(v).all()), k
    q = got["post_probs_l"].detach()
    s = q.sum(-1)
    assert torch.allclose(s, torch.ones_like(s), atol=2e-3)
    zl = got["feature"].detach()[..., 80:].reshape(B, T, dims["CL"], dims["KL"])
    assert bool(((zl == 0) | (zl == 1)).all()) and bool((zl.sum(-1) == 1).all())
    for k in MT_GRAD_IN:
        assert bool(torch.isfinite(x[k].grad).all()), k
    for k in w:
        assert w[k].grad is None or bool(torch.isfinite(w[k].grad).all()), k
    f = got["feature"].detach().cpu()
    idx_h = f[..., 32:48].reshape(B, T, dims["CH"], dims["KH"]).argmax(-1)
    idx_l = f[..., 80:].reshape(B, T, dims["CL"], dims["KL"]).argmax(-1)
    want, _, _ = oracle_mtrssm(params, inp, dims, forced=(idx_l, idx_h))
    ref = want["post_probs_l"].reshape(B, T, dims["CL"], dims["KL"])
    sure = ref.amax(-1) > 0.999
    assert float(sure.float().mean()) > 0.5  # the test really is in the extreme regime
    agree = (q.cpu().reshape(B, T, dims["CL"], dims["KL"]).argmax(-1) == ref.argmax(-1))[sure]
    assert float(agree.float().mean()) > 0.97, float(agree.float().mean())


# ---------------------------------------------------------------------------------------------------
# size-independent properties at benchmark scale, error behaviour
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", [0, 1])
def test_mtrssm_large_batch_properties(ops, precision):
    """B=4096, T=30 (bench scale): distributions normalise, samples are one-hot and equal the inverse-CDF draw of the
    kernel's own probabilities, and rolling out in two chained chunks equals one rollout (state hand-over)."""
    R, P = ops
    B, T, dims = 4096, 30, H.MT_DIMS
    params = H.make_params(H.MT_SHAPES)
    inp = cuda(H.mtrssm_inputs(B, T, dims))
    w = P.mtrssm_weight_list({k: v.cuda() for k, v in params.items()})
    kw = dict(class_size_l=4, class_size_h=2, l_tau=2.0, h_tau=4.0, precision=precision)
    full = R.mtrssm_rollout(w, **inp, **kw)
    for k in ("prior_probs_h", "prior_probs_l", "post_probs_h", "post_probs_l"):
        s = full[k].sum(-1)
        assert torch.allclose(s, torch.ones_like(s), atol=1e-5), k
    zl = full["feature"][..., 80:].reshape(B, T, 4, 4)
    assert bool(((zl == 0) | (zl == 1)).all()) and bool((zl.sum(-1) == 1).all())
    idx = O.inverse_cdf_index(full["post_probs_l"].cpu(), inp["u_post_l"].cpu())
    margin = O.cdf_margin(full["post_probs_l"].cpu(), inp["u_post_l"].cpu())
    assert bool(((idx == zl.argmax(-1).cpu()) | (margin < 1e-5)).all())
    # chained chunks
    t0 = 13
    cut = lambda d, a, b: {k: (v[:, a:b].contiguous() if v.dim() == 3 and v.shape[1] == T else v) for k, v in d.items()}  # noqa: E731
    first = R.mtrssm_rollout(w, **cut(inp, 0, t0), **kw)
    f = first["feature"][:, -1]
    nxt = cut(inp, t0, T)
    nxt.update(deter_h0=f[:, :32], stoch_h0=f[:, 32:48], deter_l0=f[:, 48:80], stoch_l0=f[:, 80:],
               hidden_h0=first["hidden_h"][:, -1], hidden_l0=first["hidden_l"][:, -1])
    second = R.mtrssm_rollout(w, **nxt, **kw)
    assert torch.equal(torch.cat([first["feature"], second["feature"]], 1), full["feature"])
    assert torch.equal(torch.cat([first["kl_l"], second["kl_l"]], 1), full["kl_l"])


def test_unsupported_sizes_and_devices_fail_loudly(ops):
    R, P = ops
    params = H.make_params(H.MR_SHAPES)
    inp = H.mrssm_inputs(4, 3)
    w = P.mrssm_weight_list({k: v.cuda() for k, v in params.items()})
    bad = cuda(inp)
    bad["embed_a"] = torch.randn(4, 3, 32, device="cuda")
    with pytest.raises(RuntimeError, match="supports"):
        R.mrssm_rollout(w, **bad)
    with pytest.raises(RuntimeError):  # CPU tensors: no fallback
        R.mrssm_rollout(P.mrssm_weight_list(params), **inp)
    odd = cuda(H.mrssm_inputs(4, 3))
    odd["actions"] = torch.randn(4, 3, 5, device="cuda")
    with pytest.raises(RuntimeError):
        R.mrssm_rollout(w, **odd)


# ---------------------------------------------------------------------------------------------------
# pre-multiplied first-layer partials (obs_projected) -- forward and every gradient
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T", [(48, 8), (300, 30)])
def test_mtrssm_obs_projected_matches_oracle_and_plain_path(ops, B, T):
    """`obs_projected=True` (the embedding half of the modality heads' first layer hoisted into one GEMM before the loop, mopoe_mmtrssm/
    core.py:259-260 reassociated) against the fp32 oracle, teacher-forced on the kernel's draws: states / probabilities within the
    stated bf16 tolerance, every gradient -- including d embed and the FULL d W1 (kernel part W1[:, :32] + GEMM part W1[:, 32:]) --
    within 2e-2 of scale.  fp32 / two-kernel policies must refuse the mode."""
    R, P = ops
    from multimodal_mtrssm_b200 import _lib

    dims = H.MT_DIMS
    params = H.make_params(H.MT_SHAPES)
    inp = H.mtrssm_inputs(B, T, dims)
    inp["u_prior_l"] = inp["u_prior_h"] = None
    up = {k: v for k, v in mtrssm_upstream(B, T, dims).items() if not k.startswith("prior_stoch")}
    w = {k: v.cuda().requires_grad_(True) for k, v in params.items()}
    x = cuda(inp)
    for k in MT_GRAD_IN:
        x[k] = x[k].requires_grad_(True)
    xin = dict(x)
    xin["embed_a"] = R.obs_projection(x["embed_a"], w["audio_representation.rnn_to_post_projector.0.weight"])
    xin["embed_v"] = R.obs_projection(x["embed_v"], w["vision_representation.rnn_to_post_projector.0.weight"])
    assert xin["embed_a"].shape == (B, T, 32)
    got = R.mtrssm_rollout(P.mtrssm_weight_list(w), class_size_l=dims["KL"], class_size_h=dims["KH"], l_tau=dims["l_tau"], h_tau=dims["h_tau"],
                           precision=_lib.PRECISION_BF16_FUSED, obs_projected=True, **xin)
    sum((got[k] * up[k].cuda()).sum() for k in up).backward()
    f = got["feature"].detach().cpu()
    idx_h = f[..., 32:48].reshape(B, T, dims["CH"], dims["KH"]).argmax(-1)
    idx_l = f[..., 80:].reshape(B, T, dims["CL"], dims["KL"]).argmax(-1)
    want, w_ref, x_ref = oracle_mtrssm(params, inp, dims, grad=True, upstream=up, forced=(idx_l, idx_h))
    rep = H.Report(f"mtrssm obs_projected B={B} T={T} vs teacher-forced oracle")
    rep.check("deter_l", f[..., 48:80], want["deter_l"], rtol=0, atol=3e-2)
    for k in ("hidden_l", "post_probs_l", "post_probs_h", "prior_probs_l"):
        rep.check(k, got[k], want[k], rtol=0, atol=3e-2)
    for k in MT_GRAD_IN:
        rep.check("d " + k, x[k].grad, x_ref[k].grad, rtol=0, atol=2e-2 * float(x_ref[k].grad.abs().max()))
    for k in w:
        rep.check("d " + k.replace("rnn_to_post_projector", "post"), w[k].grad, w_ref[k].grad, rtol=0, atol=2e-2 * float(w_ref[k].grad.abs().max()))
    rep.finish()
    for prec in (_lib.PRECISION_FP32, _lib.PRECISION_BF16):
        with pytest.raises(RuntimeError, match="obs_projected"):
            R.mtrssm_rollout(P.mtrssm_weight_list(w), precision=prec, obs_projected=True,
                             **{k: (v.detach() if v is not None else None) for k, v in xin.items()})
