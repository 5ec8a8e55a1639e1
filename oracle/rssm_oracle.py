"""CPU ORACLE for the RSSM / MTRSSM latent rollout -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this file.  Nothing under `multimodal_mtrssm_b200/` imports it; the product
path fails loudly when its CUDA library is missing.

What it is: a functional fp32 PyTorch restatement of the reference's rollout with EXPLICIT noise,
returning every per-step intermediate, differentiable by autograd.  Parameters are passed as a
dict keyed by the reference's `state_dict` names, so golden fixtures and product modules plug in
directly.  Each function cites the reference lines it follows (paths relative to
`/root/reference/src/multimodal_rssm/models/`).

PINNING STATUS.  The reference ships no tests or golden vectors, and its
distribution arithmetic lives in an absent third-party package (`distribution-extension` 1.0.7,
git nomutin/distribution-extension@e150621, `uv.lock:744-746`; `torchrl` 0.10.1 `MLP`).  The oracle
is therefore pinned like this:

* PINNED against outputs of the reference's own code run in the build container
  (`tests/golden/make_golden.py` executes `/root/reference/src/.../core.py`, `networks.py`,
  `state.py`, `mopoe_*/core.py` unmodified and commits `tests/golden/*.pt`;
  `tests/test_oracle_golden.py` checks this file against them: states, losses, gradients).
* "PARITY UNPINNED" for the third-party boundary: the semantics of `MultiOneHotFactory`,
  `rsample`, `kl_divergence(use_balancing)` and `MLP` are assumptions A1..A6,
  restated in `tests/golden/ref_shims.py` and here, each a named switch below.

Assumption switches (change here if upstream source becomes available):
  A1_CLASS_AXIS_INNERMOST  logits[..., S] viewed as [..., category, class]; softmax over class.
  A2  rsample = straight-through one-hot, flattened to [..., S]; drawn by inverse CDF from a
      caller-supplied uniform: idx = min(K-1, #{k : cdf_k <= u}).
  A5_KL_BALANCE_ALPHA = 0.8: balanced KL = a*KL(sg q||p) + (1-a)*KL(q||sg p); mean over B*T.
  A6  MLP(depth=1) = Linear -> act -> Linear (indices 0 and 2); heads use ELU.
"""

from __future__ import annotations

import math
from typing import Mapping

import torch
import torch.nn.functional as F
from torch import Tensor

A1_CLASS_AXIS_INNERMOST = True
A5_KL_BALANCE_ALPHA = 0.8

Params = Mapping[str, Tensor]


# ----------------------------------------------------------------------------------------------
# third-party arithmetic (A1..A6)
# ----------------------------------------------------------------------------------------------
def mlp(params: Params, prefix: str, x: Tensor, act=F.elu) -> tuple[Tensor, Tensor]:
    """A6: `prefix.0` Linear -> act -> `prefix.2` Linear.  Returns (out, post-activation hidden)."""
    hid = act(F.linear(x, params[f"{prefix}.0.weight"], params[f"{prefix}.0.bias"]))
    return F.linear(hid, params[f"{prefix}.2.weight"], params[f"{prefix}.2.bias"]), hid


def group_probs(logits: Tensor, C: int, K: int) -> Tensor:
    """A1: MultiOneHotFactory.forward -> probs [..., C, K]."""
    assert A1_CLASS_AXIS_INNERMOST
    return torch.softmax(logits.reshape(*logits.shape[:-1], C, K), dim=-1)


def inverse_cdf_index(probs: Tensor, u: Tensor) -> Tensor:
    """A2: idx = min(K-1, #{k : cdf_k <= u}); probs [..., C, K], u [..., C]."""
    cdf = probs.detach().cumsum(-1)
    return (cdf <= u.unsqueeze(-1)).sum(-1).clamp(max=probs.shape[-1] - 1)


def cdf_margin(probs: Tensor, u: Tensor) -> Tensor:
    """Distance of u to the nearest interior CDF boundary (tests use it to avoid knife-edge draws)."""
    cdf = probs.detach().cumsum(-1)[..., :-1]
    return (cdf - u.unsqueeze(-1)).abs().amin(-1)


def sample_st(probs: Tensor, u: Tensor | None, idx: Tensor | None = None) -> tuple[Tensor, Tensor]:
    """A2: straight-through one-hot sample flattened to [..., S]; returns (sample, idx)."""
    if idx is None:
        idx = inverse_cdf_index(probs, u)
    onehot = F.one_hot(idx, probs.shape[-1]).to(probs.dtype)
    return (onehot + probs - probs.detach()).flatten(-2), idx


def kl_per_sample(q: Tensor, p: Tensor, use_balancing: bool) -> Tensor:
    """A4+A5 before the mean: KL summed over groups, per batch element; q, p probs [..., C, K].

    Follows torch.distributions' OneHotCategorical KL (`probs * (logits_q - logits_p)`), with the
    DreamerV2 balancing split when `use_balancing`.
    """

    def kl(a: Tensor, b: Tensor) -> Tensor:
        eps = torch.finfo(a.dtype).eps
        la = a.clamp(eps, 1 - eps).log()
        lb = b.clamp(eps, 1 - eps).log()
        return (a * (la - lb)).sum(-1).sum(-1)

    if not use_balancing:
        return kl(q, p)
    a = A5_KL_BALANCE_ALPHA
    return a * kl(q.detach(), p) + (1 - a) * kl(q, p.detach())


def mopoe_fuse(audio_logits: Tensor, vision_logits: Tensor) -> Tensor:
    """MoPoE fusion, mrssm/mopoe_mrssm/core.py:241-243 + :135-154 (same at mopoe_mmtrssm/core.py:436-452).

    log-softmax over the FLAT S axis, PoE = sum (not renormalised), MoE = logsumexp of the three
    experts {audio, vision, audio+vision} each weighted 1/3.  Returns mixed "logits" [..., S].
    """
    la = F.log_softmax(audio_logits, dim=-1)
    lv = F.log_softmax(vision_logits, dim=-1)
    fused = la + lv
    log_w = math.log(1.0 / 3.0)
    return torch.logsumexp(torch.stack([log_w + la, log_w + lv, log_w + fused], dim=-2), dim=-2)


# ----------------------------------------------------------------------------------------------
# MoPoE-MRSSM
# ----------------------------------------------------------------------------------------------
def gru_cell(params: Params, x: Tensor, h: Tensor) -> tuple[Tensor, dict[str, Tensor]]:
    """nn.GRUCell (networks.py:126-129,170): gate order r,z,n; n = tanh(i_n + r*(W_hn h + b_hn))."""
    gi = F.linear(x, params["transition.rnn_cell.weight_ih"], params["transition.rnn_cell.bias_ih"])
    gh = F.linear(h, params["transition.rnn_cell.weight_hh"], params["transition.rnn_cell.bias_hh"])
    i_r, i_z, i_n = gi.chunk(3, -1)
    h_r, h_z, h_n = gh.chunk(3, -1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1 - z) * n + z * h, {"r": r, "z": z, "n": n, "h_n": h_n}


def mrssm_transition(params: Params, action: Tensor, deter: Tensor, stoch: Tensor, C: int, K: int):
    """Transition.forward, networks.py:151-173 (without the State construction/sample)."""
    x = torch.cat([action, stoch.flatten(1) if stoch.dim() >= 3 else stoch], dim=-1)  # :164-168
    x2, asp_hid = mlp(params, "transition.action_state_projector", x)  # :169
    deter, gates = gru_cell(params, x2, deter)  # :170
    prior_logits, prior_hid = mlp(params, "transition.rnn_to_prior_projector", deter)  # :171
    inter = {"asp_hid": asp_hid, "x2": x2, "prior_hid": prior_hid, "prior_logits": prior_logits, **gates}
    return deter, group_probs(prior_logits, C, K), inter  # :172


def mrssm_initial_stoch(params: Params, deter: Tensor, u: Tensor, C: int, K: int):
    """BaseRSSM.initial_state after init_proj, core.py:133-135: prior head on deter -> factory -> sample."""
    logits, _ = mlp(params, "transition.rnn_to_prior_projector", deter)
    probs = group_probs(logits, C, K)
    stoch, _ = sample_st(probs, u)
    return probs, stoch


def mrssm_rollout(
    params: Params,
    *,
    actions: Tensor,
    embed_a: Tensor,
    embed_v: Tensor,
    h0: Tensor,
    z0: Tensor,
    u_post: Tensor,
    u_prior: Tensor | None,
    C: int,
    K: int,
    forced_post_idx: Tensor | None = None,
    unimodal: bool = False,
) -> dict[str, Tensor]:
    """MoPoE_MRSSM.rollout_representation, mrssm/mopoe_mrssm/core.py:184-260, on encoder outputs.

    `unimodal=True` is BaseRSSM.rollout_representation, core.py:137-168: `posterior = representation(obs_embed[:, t], prior)`
    = Representation.forward (networks.py:70-84): factory(MLP([deter ; embed])) of the ONE head stored under
    `audio_representation.*` (the reference aliases `representation` to it), no fusion; embed_v is ignored.

    actions [B,T,A], embed_* [B,T,E], h0 [B,D], z0 [B,S], u_* [B,T,C].  The per-modality posterior
    samples the reference draws and discards (:83 -> state.py:17) are not drawn.  `forced_post_idx`
    [B,T,C] teacher-forces the posterior draw (used for reduced-precision parity).
    """
    T = actions.shape[1]
    deter, stoch = h0, z0
    keys = (
        "deter", "prior_probs", "prior_stoch", "prior_idx", "post_probs", "post_stoch", "post_idx",
        "audio_logits", "vision_logits", "mixed_logits", "audio_hid", "vision_hid",
        "asp_hid", "x2", "prior_hid", "prior_logits", "r", "z", "n", "h_n", "post_margin",
    )
    out: dict[str, list[Tensor]] = {k: [] for k in keys}
    for t in range(T):
        deter, prior_probs, inter = mrssm_transition(params, actions[:, t], deter, stoch, C, K)  # :222
        la, a_hid = mlp(params, "audio_representation.rnn_to_post_projector", torch.cat([deter, embed_a[:, t]], -1))  # :80-81
        lv, v_hid = mlp(params, "vision_representation.rnn_to_post_projector", torch.cat([deter, embed_v[:, t]], -1))
        mixed = la if unimodal else mopoe_fuse(la, lv)  # :241-251 / networks.py:82-83
        post_probs = group_probs(mixed, C, K)  # :161
        forced = None if forced_post_idx is None else forced_post_idx[:, t]
        stoch, post_idx = sample_st(post_probs, u_post[:, t], forced)  # :163 -> state.py:17
        if u_prior is not None:
            prior_stoch, prior_idx = sample_st(prior_probs, u_prior[:, t])  # networks.py:173 -> state.py:17
        else:
            prior_stoch, prior_idx = torch.zeros_like(stoch), torch.zeros_like(post_idx)
        step = {
            "deter": deter, "prior_probs": prior_probs, "prior_stoch": prior_stoch, "prior_idx": prior_idx,
            "post_probs": post_probs, "post_stoch": stoch, "post_idx": post_idx,
            "audio_logits": la, "vision_logits": lv, "mixed_logits": mixed, "audio_hid": a_hid, "vision_hid": v_hid,
            "post_margin": cdf_margin(post_probs, u_post[:, t]), **inter,
        }
        for k in keys:
            out[k].append(step[k])
    res = {k: torch.stack(v, dim=1) for k, v in out.items()}  # stack_states, state.py:121-135
    res["post_feature"] = torch.cat([res["deter"], res["post_stoch"]], -1)  # state.py:18
    return res


def mrssm_imagine(params: Params, *, actions: Tensor, h0: Tensor, z0: Tensor, u: Tensor, C: int, K: int):
    """BaseRSSM.rollout_transition, core.py:170-185: the prior's own sample is fed back."""
    deter, stoch = h0, z0
    out: dict[str, list[Tensor]] = {"deter": [], "probs": [], "stoch": [], "idx": []}
    for t in range(actions.shape[1]):
        deter, probs, _ = mrssm_transition(params, actions[:, t], deter, stoch, C, K)
        stoch, idx = sample_st(probs, u[:, t])
        for k, v in (("deter", deter), ("probs", probs), ("stoch", stoch), ("idx", idx)):
            out[k].append(v)
    return {k: torch.stack(v, dim=1) for k, v in out.items()}


def mrssm_kl(res: Mapping[str, Tensor], *, kl_coeff: float, use_balancing: bool) -> Tensor:
    """core.py:212-216: kl_divergence(q=post.independent(1), p=prior.independent(1)).mul(kl_coeff)."""
    return kl_per_sample(res["post_probs"], res["prior_probs"], use_balancing).mean() * kl_coeff


# ----------------------------------------------------------------------------------------------
# MoPoE-MMTRSSM
# ----------------------------------------------------------------------------------------------
def mtrnn(params: Params, prefix: str, x: Tensor, prev_d: Tensor, hidden: Tensor, tau: float):
    """MTRNN._compute_mtrnn, mmtrssm/mopoe_mmtrssm/core.py:59-60 (hidden passed in/out, not a module attr)."""
    pre = F.linear(prev_d, params[f"{prefix}._d2h.weight"], params[f"{prefix}._d2h.bias"]) + F.linear(
        x, params[f"{prefix}._input2h.weight"], params[f"{prefix}._input2h.bias"]
    )
    hidden = (1 - 1 / tau) * hidden + pre / tau
    return torch.tanh(hidden), hidden


def mtrssm_prior_step(params: Params, action, d_l, d_h, u_l, u_h, z_l, z_h, dims):
    """_compute_lower_prior (:263-287) + the h_rnn/h_prior half of _compute_higher_prior_posterior (:309-312)."""
    x_l = torch.cat([action, z_l, z_h], dim=-1)  # :283
    d_l, u_l = mtrnn(params, "l_rnn", x_l, d_l, u_l, dims["l_tau"])  # :284
    lp_logits, lp_hid = mlp(params, "l_prior", d_l)  # :285
    d_h, u_h = mtrnn(params, "h_rnn", z_h, d_h, u_h, dims["h_tau"])  # :310
    hp_logits, hp_hid = mlp(params, "h_prior", d_h)  # :311
    return d_l, u_l, d_h, u_h, lp_logits, hp_logits, lp_hid, hp_hid


def mtrssm_initial_stoch(params: Params, higher: Tensor, lower: Tensor, u_h: Tensor, u_l: Tensor, dims):
    """MoPoE_MMTRSSM.initial_state after init_proj, mopoe_mmtrssm/core.py:349-361 (h sampled first)."""
    p_h = group_probs(mlp(params, "h_prior", higher)[0], int(dims["CH"]), int(dims["KH"]))
    p_l = group_probs(mlp(params, "l_prior", lower)[0], int(dims["CL"]), int(dims["KL"]))
    return p_h, p_l, sample_st(p_h, u_h)[0], sample_st(p_l, u_l)[0]


def mtrssm_rollout(
    params: Params,
    *,
    actions: Tensor,
    embed_a: Tensor,
    embed_v: Tensor,
    deter_h0: Tensor,
    deter_l0: Tensor,
    hidden_h0: Tensor,
    hidden_l0: Tensor,
    stoch_h0: Tensor,
    stoch_l0: Tensor,
    u_post_l: Tensor,
    u_post_h: Tensor,
    u_prior_l: Tensor | None,
    u_prior_h: Tensor | None,
    dims: Mapping[str, float],
    forced_idx_l: Tensor | None = None,
    forced_idx_h: Tensor | None = None,
) -> dict[str, Tensor]:
    """MoPoE_MMTRSSM.rollout_representation, mmtrssm/mopoe_mmtrssm/core.py:364-494, on encoder outputs.

    dims: CL,KL (l_dist), CH,KH (h_dist), l_tau, h_tau.  `hidden_*0` seeds MTRNN.hidden
    (`_set_prev_hiddens`, :206-239, called at :400).  `l_posterior` is never used (:405-490).
    """
    CL, KL, CH, KH = (int(dims[k]) for k in ("CL", "KL", "CH", "KH"))
    d_h, d_l, u_h, u_l, z_h, z_l = deter_h0, deter_l0, hidden_h0, hidden_l0, stoch_h0, stoch_l0
    keys = (
        "deter_h", "deter_l", "hidden_h", "hidden_l", "prior_probs_h", "prior_probs_l", "post_probs_h",
        "post_probs_l", "post_stoch_h", "post_stoch_l", "post_idx_h", "post_idx_l", "prior_stoch_h",
        "prior_stoch_l", "audio_logits", "vision_logits", "audio_hid", "vision_hid", "lp_hid", "hp_hid",
        "hq_hid", "margin_l", "margin_h",
    )
    out: dict[str, list[Tensor]] = {k: [] for k in keys}
    for t in range(actions.shape[1]):
        d_l, u_l, d_h, u_h, lp_logits, hp_logits, lp_hid, hp_hid = mtrssm_prior_step(
            params, actions[:, t], d_l, d_h, u_l, u_h, z_l, z_h, dims
        )
        la, a_hid = mlp(params, "audio_representation.rnn_to_post_projector", torch.cat([d_l, embed_a[:, t]], -1))  # :259-260
        lv, v_hid = mlp(params, "vision_representation.rnn_to_post_projector", torch.cat([d_l, embed_v[:, t]], -1))
        post_l = group_probs(mopoe_fuse(la, lv), CL, KL)  # :436-455
        z_l, idx_l = sample_st(post_l, u_post_l[:, t], None if forced_idx_l is None else forced_idx_l[:, t])  # :456
        hq_logits, hq_hid = mlp(params, "h_posterior", torch.cat([d_l, d_h], -1))  # :315-316
        post_h = group_probs(hq_logits, CH, KH)  # :317
        z_h, idx_h = sample_st(post_h, u_post_h[:, t], None if forced_idx_h is None else forced_idx_h[:, t])  # :464
        prior_l, prior_h = group_probs(lp_logits, CL, KL), group_probs(hp_logits, CH, KH)
        if u_prior_l is not None:  # prior MTState ctor samples h then l (:467-474 -> state.py:48-49)
            pz_h, _ = sample_st(prior_h, u_prior_h[:, t])
            pz_l, _ = sample_st(prior_l, u_prior_l[:, t])
        else:
            pz_h, pz_l = torch.zeros_like(z_h), torch.zeros_like(z_l)
        step = {
            "deter_h": d_h, "deter_l": d_l, "hidden_h": u_h, "hidden_l": u_l,
            "prior_probs_h": prior_h, "prior_probs_l": prior_l, "post_probs_h": post_h, "post_probs_l": post_l,
            "post_stoch_h": z_h, "post_stoch_l": z_l, "post_idx_h": idx_h, "post_idx_l": idx_l,
            "prior_stoch_h": pz_h, "prior_stoch_l": pz_l, "audio_logits": la, "vision_logits": lv,
            "audio_hid": a_hid, "vision_hid": v_hid, "lp_hid": lp_hid, "hp_hid": hp_hid, "hq_hid": hq_hid,
            "margin_l": cdf_margin(post_l, u_post_l[:, t]), "margin_h": cdf_margin(post_h, u_post_h[:, t]),
        }
        for k in keys:
            out[k].append(step[k])
    res = {k: torch.stack(v, dim=1) for k, v in out.items()}  # stack_mtstates, mmtrssm/state.py:184-215
    res["post_feature"] = torch.cat(
        [res["deter_h"], res["post_stoch_h"], res["deter_l"], res["post_stoch_l"]], -1
    )  # mmtrssm/state.py:51
    return res


def mtrssm_imagine(
    params: Params, *, actions, deter_h0, deter_l0, hidden_h0, hidden_l0, stoch_h0, stoch_l0, u_l, u_h, dims
):
    """MoPoE_MMTRSSM.rollout_transition, mmtrssm/mopoe_mmtrssm/core.py:496-544 (prior samples fed back)."""
    CL, KL, CH, KH = (int(dims[k]) for k in ("CL", "KL", "CH", "KH"))
    d_h, d_l, hid_h, hid_l, z_h, z_l = deter_h0, deter_l0, hidden_h0, hidden_l0, stoch_h0, stoch_l0
    keys = ("deter_h", "deter_l", "hidden_h", "hidden_l", "probs_h", "probs_l", "stoch_h", "stoch_l")
    out: dict[str, list[Tensor]] = {k: [] for k in keys}
    for t in range(actions.shape[1]):
        d_l, hid_l, d_h, hid_h, lp_logits, hp_logits, _, _ = mtrssm_prior_step(
            params, actions[:, t], d_l, d_h, hid_l, hid_h, z_l, z_h, dims
        )
        p_h, p_l = group_probs(hp_logits, CH, KH), group_probs(lp_logits, CL, KL)
        z_h, _ = sample_st(p_h, u_h[:, t])
        z_l, _ = sample_st(p_l, u_l[:, t])
        for k, v in zip(keys, (d_h, d_l, hid_h, hid_l, p_h, p_l, z_h, z_l)):
            out[k].append(v)
    return {k: torch.stack(v, dim=1) for k, v in out.items()}


def mtrssm_kl(res: Mapping[str, Tensor], *, kl_coeff: float, w_kl_h: float, use_balancing: bool):
    """mmtrssm/mopoe_mmtrssm/core.py:589-600: (kl_l * kl_coeff, kl_h * kl_coeff * w_kl_h)."""
    kl_l = kl_per_sample(res["post_probs_l"], res["prior_probs_l"], use_balancing).mean() * kl_coeff
    kl_h = kl_per_sample(res["post_probs_h"], res["prior_probs_h"], use_balancing).mean() * (kl_coeff * w_kl_h)
    return kl_l, kl_h


# ----------------------------------------------------------------------------------------------
# "what the reference does today": literal per-step structure, for CPU-baseline timing only
# ----------------------------------------------------------------------------------------------
def mrssm_rollout_literal(params: Params, *, actions, embed_a, embed_v, h0, z0, C: int, K: int):
    """Same arithmetic as `mrssm_rollout` but keeping the reference's wasted work: four global-RNG
    categorical draws per step (prior, audio, vision, mixed; state.py:17), the duplicated
    log-softmaxes (mopoe_mrssm/core.py:241-242 and :136-137), a fresh log(1/3) tensor per step
    (:140-141), per-step feature cats and T-way stacks.  Used by bench.py's reference arm."""

    def draw(p: Tensor) -> Tensor:
        idx = torch.multinomial(p.detach().reshape(-1, K), 1).reshape(p.shape[:-1])
        return (F.one_hot(idx, K).to(p.dtype) + p - p.detach()).flatten(-2)

    deter, stoch = h0, z0
    feats, priors, posts = [], [], []
    for t in range(actions.shape[1]):
        deter, prior_probs, _ = mrssm_transition(params, actions[:, t], deter, stoch, C, K)
        _ = torch.cat([deter, draw(prior_probs)], -1)
        la, _h = mlp(params, "audio_representation.rnn_to_post_projector", torch.cat([deter, embed_a[:, t]], -1))
        _ = torch.cat([deter, draw(group_probs(la, C, K))], -1)
        lv, _h = mlp(params, "vision_representation.rnn_to_post_projector", torch.cat([deter, embed_v[:, t]], -1))
        _ = torch.cat([deter, draw(group_probs(lv, C, K))], -1)
        fused = F.log_softmax(la, -1) + F.log_softmax(lv, -1)
        a2, v2 = F.log_softmax(la, -1), F.log_softmax(lv, -1)
        log_w = torch.log(torch.tensor(1.0 / 3.0, dtype=a2.dtype))
        mixed = torch.logsumexp(torch.stack([log_w + a2, log_w + v2, log_w + fused], dim=-2), dim=-2)
        post_probs = group_probs(mixed, C, K)
        stoch = draw(post_probs)
        feats.append(torch.cat([deter, stoch], -1))
        priors.append(prior_probs)
        posts.append(post_probs)
    return {"post_feature": torch.stack(feats, 1), "prior_probs": torch.stack(priors, 1), "post_probs": torch.stack(posts, 1)}


def mtrssm_rollout_literal(params: Params, *, actions, embed_a, embed_v, deter_h0, deter_l0, hidden_h0, hidden_l0, stoch_h0, stoch_l0, dims):
    """Literal-structure MMTRSSM rollout (mopoe_mmtrssm/core.py:405-490) with global-RNG draws."""
    CL, KL, CH, KH = (int(dims[k]) for k in ("CL", "KL", "CH", "KH"))

    def draw(p: Tensor) -> Tensor:
        k = p.shape[-1]
        idx = torch.multinomial(p.detach().reshape(-1, k), 1).reshape(p.shape[:-1])
        return (F.one_hot(idx, k).to(p.dtype) + p - p.detach()).flatten(-2)

    d_h, d_l, u_h, u_l, z_h, z_l = deter_h0, deter_l0, hidden_h0, hidden_l0, stoch_h0, stoch_l0
    feats, pl, ph, ql, qh = [], [], [], [], []
    for t in range(actions.shape[1]):
        d_l, u_l, d_h, u_h, lp_logits, hp_logits, _, _ = mtrssm_prior_step(params, actions[:, t], d_l, d_h, u_l, u_h, z_l, z_h, dims)
        la, _h = mlp(params, "audio_representation.rnn_to_post_projector", torch.cat([d_l, embed_a[:, t]], -1))
        lv, _h = mlp(params, "vision_representation.rnn_to_post_projector", torch.cat([d_l, embed_v[:, t]], -1))
        a2, v2 = F.log_softmax(la, -1), F.log_softmax(lv, -1)
        log_w = torch.log(torch.tensor(1.0 / 3.0, dtype=a2.dtype))
        mixed = torch.logsumexp(torch.stack([log_w + a2, log_w + v2, log_w + (a2 + v2)], dim=-2), dim=-2)
        post_l = group_probs(mixed, CL, KL)
        z_l = draw(post_l)
        hq_logits, _h = mlp(params, "h_posterior", torch.cat([d_l, d_h], -1))
        post_h = group_probs(hq_logits, CH, KH)
        z_h = draw(post_h)
        prior_l, prior_h = group_probs(lp_logits, CL, KL), group_probs(hp_logits, CH, KH)
        _ = torch.cat([d_h, draw(prior_h), d_l, draw(prior_l)], -1)  # prior MTState ctor
        feats.append(torch.cat([d_h, z_h, d_l, z_l], -1))
        pl.append(prior_l), ph.append(prior_h), ql.append(post_l), qh.append(post_h)
    st = lambda xs: torch.stack(xs, 1)  # noqa: E731
    return {"post_feature": st(feats), "prior_probs_l": st(pl), "prior_probs_h": st(ph), "post_probs_l": st(ql), "post_probs_h": st(qh)}


# ---------------------------------------------------------------------------------------------------------------------
# reconstruction likelihood
# ---------------------------------------------------------------------------------------------------------------------
def likelihood(prediction: Tensor, target: Tensor, event_ndims: int, scale: float = 1.0) -> Tensor:
    """objective.py:21-23, literally: -Independent(Normal(prediction, scale), event_ndims).log_prob(target).mean()
    (torch.distributions is the arithmetic the reference itself calls, so this row of the oracle is pinned by torch)."""
    import torch.distributions as td

    return -td.Independent(td.Normal(prediction, scale), event_ndims).log_prob(target).mean()
