"""GPU parity of the WIDE MoPoE-MRSSM rollout (deterministic_size = hidden_size = D up to 512: BASELINE.json cfg3) against
the CPU oracle, through the torch custom ops -> C ABI -> persistent tcgen05 kernels.

The wide family computes on the bf16 tensor-core path only.  Stated tolerances (written at each check): 3e-2 absolute on the
O(1) states / probabilities of a rollout that is teacher-forced on the kernel's own categorical draws (bf16 operands with
fp32 accumulation over K = D <= 512, error compounding over the recurrence), 3e-2 of each gradient tensor's scale.
"""

from __future__ import annotations

import pytest
import torch

from oracle import rssm_oracle as O
from tests import helpers as H
from tests.test_rollout_gpu import cuda, oracle_mrssm, run_mrssm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from multimodal_mtrssm_b200 import params as P
    from multimodal_mtrssm_b200 import rollout_ops as R

    return R, P


def wide_case(D, B, T, K, gain=2.0):
    C = 16 // K
    params = H.make_params(H.mr_shapes(D), gain=gain)
    inp = H.mrssm_inputs(B, T, C, K, D=D)
    return params, inp, C


@pytest.mark.parametrize("D,B,T,K", [(512, 200, 6, 4), (128, 128, 5, 2), (384, 33, 4, 16), (512, 1, 1, 4), (256, 300, 3, 8)])
def test_wide_forward_bf16_teacher_forced(ops, D, B, T, K):
    R, P = ops
    params, inp, C = wide_case(D, B, T, K)
    got, _, _ = run_mrssm(R, P, params, inp, K, precision=1)
    onehot = got["feature"][..., D:].cpu().reshape(B, T, C, K)
    assert bool(((onehot == 0) | (onehot == 1)).all()) and bool((onehot.sum(-1) == 1).all())
    idx = onehot.argmax(-1)
    want, _, _ = oracle_mrssm(params, inp, C, K, forced=idx)
    rep = H.Report(f"wide mrssm fwd bf16 D={D} B={B} T={T} K={K} (teacher-forced on the kernel's draws)")
    rep.check("deter", got["feature"][..., :D], want["deter"], rtol=0, atol=3e-2)
    rep.check("prior_probs", got["prior_probs"], want["prior_probs"], rtol=0, atol=3e-2)
    rep.check("post_probs", got["post_probs"], want["post_probs"], rtol=0, atol=3e-2)
    rep.check("kl", got["kl"], want["kl"], rtol=5e-2, atol=5e-2)
    rep.finish()
    # the kernel's draws are the inverse-CDF draws of ITS OWN probabilities (posterior and prior)
    self_idx = O.inverse_cdf_index(got["post_probs"].cpu(), inp["u_post"])
    margin = O.cdf_margin(got["post_probs"].cpu(), inp["u_post"])
    assert bool(((self_idx == idx) | (margin < 1e-5)).all())
    pidx = got["prior_stoch"].cpu().reshape(B, T, C, K).argmax(-1)
    self_p = O.inverse_cdf_index(got["prior_probs"].cpu(), inp["u_prior"])
    margin_p = O.cdf_margin(got["prior_probs"].cpu(), inp["u_prior"])
    assert bool(((self_p == pidx) | (margin_p < 1e-5)).all())
    for k in ("prior_probs", "post_probs"):
        s = got[k].sum(-1)
        assert torch.allclose(s, torch.ones_like(s), atol=1e-3), k


def wide_upstream(B, T, C, K, D, seed=7):
    g = torch.Generator().manual_seed(seed)
    return {
        "feature": torch.randn(B, T, D + 16, generator=g), "kl": torch.randn(B, T, generator=g),
        "post_probs": torch.randn(B, T, C, K, generator=g), "prior_probs": torch.randn(B, T, C, K, generator=g),
    }


@pytest.mark.parametrize("D,B,T,K,balancing", [(512, 200, 5, 4, True), (128, 96, 6, 2, False), (256, 130, 1, 8, True)])
def test_wide_backward_bf16_vs_oracle(ops, D, B, T, K, balancing):
    """All gradients (inputs, initial state, every weight) of the wide bf16 path against the fp32 oracle, teacher-forced on the
    kernel's own draws.  Stated bf16 tolerance: 3e-2 of each gradient tensor's scale (max |grad|)."""
    R, P = ops
    params, inp, C = wide_case(D, B, T, K)
    inp["u_prior"] = None  # the prior sample is not part of the training loss
    up = wide_upstream(B, T, C, K, D)
    got, w, x = run_mrssm(R, P, params, inp, K, precision=1, grad=True, upstream=up, use_balancing=balancing)
    idx = got["feature"][..., D:].detach().cpu().reshape(B, T, C, K).argmax(-1)
    _, w_ref, x_ref = oracle_mrssm(params, inp, C, K, grad=True, upstream=up, use_balancing=balancing, forced=idx)
    rep = H.Report(f"wide mrssm bwd bf16 D={D} B={B} T={T} K={K} vs teacher-forced oracle")
    for k in ("actions", "embed_a", "embed_v", "h0", "z0"):
        rep.check("d " + k, x[k].grad, x_ref[k].grad, rtol=0, atol=3e-2 * float(x_ref[k].grad.abs().max()))
    for k in w:
        rep.check("d " + k.replace("rnn_to_post_projector", "post"), w[k].grad, w_ref[k].grad, rtol=0,
                  atol=3e-2 * float(w_ref[k].grad.abs().max()))
    rep.finish()


def test_wide_multiple_launch_groups_match_oracle(ops):
    """B = 1300 at hidden 512 is 11 batch blocks = two launch groups (9 + 2 blocks on 148 SMs): the group offsets of every
    buffer (records, gradient planes, narrow planes, statistics) are exercised; forward and all gradients against the oracle."""
    R, P = ops
    D, B, T, K = 512, 1300, 2, 4
    params, inp, C = wide_case(D, B, T, K)
    inp["u_prior"] = None
    up = wide_upstream(B, T, C, K, D)
    got, w, x = run_mrssm(R, P, params, inp, K, precision=1, grad=True, upstream=up)
    idx = got["feature"][..., D:].detach().cpu().reshape(B, T, C, K).argmax(-1)
    want, w_ref, x_ref = oracle_mrssm(params, inp, C, K, grad=True, upstream=up, forced=idx)
    rep = H.Report("wide mrssm, two launch groups (B=1300, hidden 512)")
    rep.check("deter", got["feature"][..., :D], want["deter"], rtol=0, atol=3e-2)
    rep.check("post_probs", got["post_probs"], want["post_probs"], rtol=0, atol=3e-2)
    for k in ("actions", "embed_a", "embed_v", "h0", "z0"):
        rep.check("d " + k, x[k].grad, x_ref[k].grad, rtol=0, atol=3e-2 * float(x_ref[k].grad.abs().max()))
    for k in w:
        rep.check("d " + k.replace("rnn_to_post_projector", "post"), w[k].grad, w_ref[k].grad, rtol=0,
                  atol=3e-2 * float(w_ref[k].grad.abs().max()))
    rep.finish()


def test_wide_cfg3_full_size_properties(ops):
    """BASELINE.json cfg3 at full size (B = 1024, T = 64, hidden 512), where the CPU oracle is too slow: size-independent
    properties.  Distributions normalise; samples are exact one-hots and equal the inverse-CDF draw of the kernel's own
    probabilities; the KL output equals the KL of the returned probabilities; rolling out in two chained chunks equals one
    rollout BIT FOR BIT (state hand-over, deterministic summation orders); gradients are finite and linear in the upstream."""
    R, P = ops
    D, B, T, K, C = 512, 1024, 64, 4, 4
    params = {k: v.cuda() for k, v in H.make_params(H.mr_shapes(D), gain=1.0).items()}
    inp = cuda(H.mrssm_inputs(B, T, C, K, D=D))
    inp["u_prior"] = None
    w = P.mrssm_weight_list(params)
    with torch.no_grad():
        full = R.mrssm_rollout(w, class_size=K, precision=1, **inp)
    for k in ("prior_probs", "post_probs"):
        s = full[k].sum(-1)
        assert torch.allclose(s, torch.ones_like(s), atol=1e-3), k
    z = full["feature"][..., D:].reshape(B, T, C, K)
    assert bool(((z == 0) | (z == 1)).all()) and bool((z.sum(-1) == 1).all())
    idx = O.inverse_cdf_index(full["post_probs"].cpu(), inp["u_post"].cpu())
    margin = O.cdf_margin(full["post_probs"].cpu(), inp["u_post"].cpu())
    assert bool(((idx == z.argmax(-1).cpu()) | (margin < 1e-5)).all())
    kl = O.kl_per_sample(full["post_probs"].cpu(), full["prior_probs"].cpu(), False)
    torch.testing.assert_close(full["kl"].cpu(), kl, rtol=2e-3, atol=2e-3)
    assert bool(torch.isfinite(full["feature"]).all()) and float(full["feature"][..., :D].abs().max()) <= 1.0 + 1e-3 + float(inp["h0"].abs().max())
    # chained chunks
    t0 = 23
    cut = lambda d, a, b: {k: (v[:, a:b].contiguous() if v is not None and v.dim() == 3 and v.shape[1] == T else v) for k, v in d.items()}  # noqa: E731
    with torch.no_grad():
        first = R.mrssm_rollout(w, class_size=K, precision=1, **cut(inp, 0, t0))
        nxt = cut(inp, t0, T)
        nxt.update(h0=first["feature"][:, -1, :D].contiguous(), z0=first["feature"][:, -1, D:].contiguous())
        second = R.mrssm_rollout(w, class_size=K, precision=1, **nxt)
    assert torch.equal(torch.cat([first["feature"], second["feature"]], 1), full["feature"])
    assert torch.equal(torch.cat([first["kl"], second["kl"]], 1), full["kl"])
    # gradients: finite, and linear in the upstream gradient (2x upstream -> 2x gradients up to bf16 rounding of the planes)
    wg = [t.clone().requires_grad_(True) for t in w]
    up = torch.randn(B, T, D + 16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    out = R.mrssm_rollout(wg, class_size=K, precision=1, **inp)
    g1 = torch.autograd.grad((out["feature"] * up).sum(), wg, retain_graph=True)
    g2 = torch.autograd.grad((out["feature"] * (2 * up)).sum(), wg)
    for a, b in zip(g1, g2):
        assert bool(torch.isfinite(a).all())
        torch.testing.assert_close(b, 2 * a, rtol=0, atol=2e-2 * max(float(a.abs().max()), 1e-6))


def test_wide_family_forward_backward_replays_from_a_cuda_graph(ops):
    """The wide family's forward + backward (weight packing, persistent kernels, pre-pass, weight-gradient tile table, all the
    non-recurrent kernels) captured ONCE into a CUDA graph and replayed on new data: nothing in the path may be a pageable
    host-to-device copy or depend on host state that a later call could change (the weight-gradient tile table is a by-value
    kernel parameter).  Replays must match eager calls on the same data."""
    R, P = ops
    D, B, T, K, C = 128, 96, 3, 4, 4
    params = {k: v.cuda() for k, v in H.make_params(H.mr_shapes(D)).items()}
    w = [t.clone().requires_grad_(True) for t in P.mrssm_weight_list(params)]
    static = cuda(H.mrssm_inputs(B, T, C, K, D=D))
    static["u_prior"] = None
    up = torch.randn(B, T, D + 16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))

    def step(inp):
        out = R.mrssm_rollout(w, class_size=K, precision=1, **inp)
        b, t = out["feature"].shape[:2]
        grads = torch.autograd.grad((out["feature"] * up[:b, :t]).sum() + out["kl"].mean(), w)
        return out["feature"], grads

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step(static)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_feature, g_grads = step(static)
    for seed in (11, 12):
        fresh = cuda(H.mrssm_inputs(B, T, C, K, seed=seed, D=D))
        for k, v in static.items():
            if v is not None:
                v.copy_(fresh[k])
        # an unrelated wide call with OTHER sizes in between must not disturb the captured graph
        other = cuda(H.mrssm_inputs(40, 2, C, K, D=D))
        other["u_prior"] = None
        step(other)
        graph.replay()
        torch.cuda.synchronize()
        e_feature, e_grads = step(static)
        assert torch.equal(g_feature, e_feature)
        for a, b in zip(g_grads, e_grads):  # atomics: summation order differs between launches
            torch.testing.assert_close(a, b, rtol=0, atol=2e-3 * max(float(b.abs().max()), 1e-6))
