"""Shared test helpers: synthetic parameters/inputs and knife-edge-free uniforms."""

from __future__ import annotations

import torch

from oracle import rssm_oracle as O

MR_SHAPES = {
    "transition.action_state_projector.0.weight": (32, 22), "transition.action_state_projector.0.bias": (32,),
    "transition.action_state_projector.2.weight": (32, 32), "transition.action_state_projector.2.bias": (32,),
    "transition.rnn_cell.weight_ih": (96, 32), "transition.rnn_cell.weight_hh": (96, 32),
    "transition.rnn_cell.bias_ih": (96,), "transition.rnn_cell.bias_hh": (96,),
    "transition.rnn_to_prior_projector.0.weight": (32, 32), "transition.rnn_to_prior_projector.0.bias": (32,),
    "transition.rnn_to_prior_projector.2.weight": (16, 32), "transition.rnn_to_prior_projector.2.bias": (16,),
    "audio_representation.rnn_to_post_projector.0.weight": (32, 96), "audio_representation.rnn_to_post_projector.0.bias": (32,),
    "audio_representation.rnn_to_post_projector.2.weight": (16, 32), "audio_representation.rnn_to_post_projector.2.bias": (16,),
    "vision_representation.rnn_to_post_projector.0.weight": (32, 96), "vision_representation.rnn_to_post_projector.0.bias": (32,),
    "vision_representation.rnn_to_post_projector.2.weight": (16, 32), "vision_representation.rnn_to_post_projector.2.bias": (16,),
}



def mr_shapes(D: int, A: int = 6) -> dict:
    """MoPoE-MRSSM parameter shapes for deterministic_size = hidden_size = D (BASELINE.json cfg3: D = 512)."""
    t, a, v = "transition.", "audio_representation.rnn_to_post_projector.", "vision_representation.rnn_to_post_projector."
    return {
        t + "action_state_projector.0.weight": (D, A + 16), t + "action_state_projector.0.bias": (D,),
        t + "action_state_projector.2.weight": (D, D), t + "action_state_projector.2.bias": (D,),
        t + "rnn_cell.weight_ih": (3 * D, D), t + "rnn_cell.weight_hh": (3 * D, D),
        t + "rnn_cell.bias_ih": (3 * D,), t + "rnn_cell.bias_hh": (3 * D,),
        t + "rnn_to_prior_projector.0.weight": (D, D), t + "rnn_to_prior_projector.0.bias": (D,),
        t + "rnn_to_prior_projector.2.weight": (16, D), t + "rnn_to_prior_projector.2.bias": (16,),
        a + "0.weight": (D, D + 64), a + "0.bias": (D,), a + "2.weight": (16, D), a + "2.bias": (16,),
        v + "0.weight": (D, D + 64), v + "0.bias": (D,), v + "2.weight": (16, D), v + "2.bias": (16,),
    }


MT_SHAPES = {
    "l_rnn._d2h.weight": (32, 32), "l_rnn._d2h.bias": (32,), "l_rnn._input2h.weight": (32, 38), "l_rnn._input2h.bias": (32,),
    "h_rnn._d2h.weight": (32, 32), "h_rnn._d2h.bias": (32,), "h_rnn._input2h.weight": (32, 16), "h_rnn._input2h.bias": (32,),
    "l_prior.0.weight": (32, 32), "l_prior.0.bias": (32,), "l_prior.2.weight": (16, 32), "l_prior.2.bias": (16,),
    "h_prior.0.weight": (32, 32), "h_prior.0.bias": (32,), "h_prior.2.weight": (16, 32), "h_prior.2.bias": (16,),
    "h_posterior.0.weight": (32, 64), "h_posterior.0.bias": (32,), "h_posterior.2.weight": (16, 32), "h_posterior.2.bias": (16,),
    "audio_representation.rnn_to_post_projector.0.weight": (32, 96), "audio_representation.rnn_to_post_projector.0.bias": (32,),
    "audio_representation.rnn_to_post_projector.2.weight": (16, 32), "audio_representation.rnn_to_post_projector.2.bias": (16,),
    "vision_representation.rnn_to_post_projector.0.weight": (32, 96), "vision_representation.rnn_to_post_projector.0.bias": (32,),
    "vision_representation.rnn_to_post_projector.2.weight": (16, 32), "vision_representation.rnn_to_post_projector.2.bias": (16,),
}

MT_DIMS = dict(CL=4, KL=4, CH=8, KH=2, l_tau=2.0, h_tau=4.0)


def make_params(shapes: dict, seed: int = 42, gain: float = 2.0) -> dict[str, torch.Tensor]:
    """U(-g/sqrt(fan_in), g/sqrt(fan_in)); gain > 1 makes the categoricals peaky enough to be a real test."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, s in shapes.items():
        fan_in = s[1] if len(s) == 2 else 32
        out[k] = (torch.rand(*s, generator=g) * 2 - 1) * gain / fan_in**0.5
    return out


def onehot_draw(B: int, C: int, K: int, g: torch.Generator) -> torch.Tensor:
    idx = torch.randint(0, K, (B, C), generator=g)
    return torch.nn.functional.one_hot(idx, K).float().flatten(1)


def synth_actions(B: int, T: int, g: torch.Generator, A: int = 6) -> torch.Tensor:
    speaker = torch.randint(0, A, (B,), generator=g)
    act = torch.nn.functional.one_hot(speaker, A).float()[:, None, :].expand(B, T, A)
    return (act + 0.1 * torch.randn(B, T, A, generator=g)).contiguous()


def mrssm_inputs(B: int, T: int, C: int = 4, K: int = 4, seed: int = 1234, D: int = 32) -> dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    n = torch.Generator().manual_seed(4321)
    return {
        "actions": synth_actions(B, T, g), "embed_a": torch.randn(B, T, 64, generator=g), "embed_v": torch.randn(B, T, 64, generator=g),
        "h0": torch.randn(B, D, generator=g), "z0": onehot_draw(B, C, K, g),
        "u_post": torch.rand(B, T, C, generator=n), "u_prior": torch.rand(B, T, C, generator=n),
    }


def mtrssm_inputs(B: int, T: int, dims: dict = MT_DIMS, seed: int = 1234) -> dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    n = torch.Generator().manual_seed(4321)
    CL, KL, CH, KH = dims["CL"], dims["KL"], dims["CH"], dims["KH"]
    d_h, d_l = torch.randn(B, 32, generator=g), torch.randn(B, 32, generator=g)
    return {
        "actions": synth_actions(B, T, g), "embed_a": torch.randn(B, T, 64, generator=g), "embed_v": torch.randn(B, T, 64, generator=g),
        "deter_h0": d_h, "deter_l0": d_l, "hidden_h0": d_h.clone(), "hidden_l0": d_l.clone(),  # raw init, a18
        "stoch_h0": onehot_draw(B, CH, KH, g), "stoch_l0": onehot_draw(B, CL, KL, g),
        "u_post_l": torch.rand(B, T, CL, generator=n), "u_post_h": torch.rand(B, T, CH, generator=n),
        "u_prior_l": torch.rand(B, T, CL, generator=n), "u_prior_h": torch.rand(B, T, CH, generator=n),
    }


def _resample(u: torch.Tensor, bad: torch.Tensor, g: torch.Generator) -> int:
    n = int(bad.sum())
    if n:
        u[bad] = torch.rand(n, generator=g)
    return n


def mrssm_safe_uniforms(params, inp, C: int, K: int, eps: float, unimodal: bool = False) -> None:
    """Resample (in place) uniforms that sit within `eps` of a CDF boundary of the ORACLE trajectory, so that
    rounding-level differences between two correct implementations cannot flip a categorical draw."""
    g = torch.Generator().manual_seed(99)
    for _ in range(50):
        with torch.no_grad():
            res = O.mrssm_rollout(params, C=C, K=K, unimodal=unimodal, **inp)
            bad = res["post_margin"] < eps
            badp = O.cdf_margin(res["prior_probs"], inp["u_prior"]) < eps
        if _resample(inp["u_post"], bad, g) + _resample(inp["u_prior"], badp, g) == 0:
            return
    raise AssertionError("could not find knife-edge-free uniforms")


def mtrssm_safe_uniforms(params, inp, dims, eps: float) -> None:
    g = torch.Generator().manual_seed(99)
    for _ in range(50):
        with torch.no_grad():
            res = O.mtrssm_rollout(params, dims=dims, **inp)
            n = _resample(inp["u_post_l"], res["margin_l"] < eps, g) + _resample(inp["u_post_h"], res["margin_h"] < eps, g)
            n += _resample(inp["u_prior_l"], O.cdf_margin(res["prior_probs_l"], inp["u_prior_l"]) < eps, g)
            n += _resample(inp["u_prior_h"], O.cdf_margin(res["prior_probs_h"], inp["u_prior_h"]) < eps, g)
        if n == 0:
            return
    raise AssertionError("could not find knife-edge-free uniforms")


def golden_batch(g: dict) -> tuple[torch.Tensor, ...]:
    """The 6-tuple batch of a golden fixture.  The default.yaml-sized fixtures store the seed of make_golden.synth_batch (same
    generator calls here) plus a checksum instead of 2 x 4 MB of observations."""
    if "batch" in g:
        return tuple(g["batch"])
    B, T = g["dims"]["B"], g["dims"]["T"]
    gen = torch.Generator().manual_seed(g["batch_seed"])
    speaker = torch.randint(0, 6, (B,), generator=gen)
    act = torch.nn.functional.one_hot(speaker, 6).float()[:, None, :].expand(B, T, 6)
    act_in = act + 0.1 * torch.randn(B, T, 6, generator=gen)
    audio = torch.rand(B, T, 1, 32, 32, generator=gen) * 2 - 1
    vision = torch.rand(B, T, 1, 32, 32, generator=gen) * 2 - 1
    batch = (act_in, audio, vision, act.clone(), audio.clone(), vision.clone())
    assert [float(t.double().sum()) for t in batch] == g["batch_checksum"], "golden batch does not regenerate bit-exactly"
    assert torch.equal(batch[0], g["inputs"]["actions"])
    return batch


MRSSM_GOLDEN = ("mrssm_default.pt", "mrssm_cfg1.pt")   # B=5,T=7 and default.yaml's own B=8,T=30 (BASELINE.json configs[0])
MTRSSM_GOLDEN = ("mtrssm_default.pt", "mtrssm_cfg2.pt")  # B=5,T=7 and default.yaml's own B=8,T=30 (BASELINE.json configs[1])


class Report:
    """Collects per-tensor errors so one GPU run shows every mismatch, then asserts."""

    def __init__(self, title: str) -> None:
        self.title, self.rows, self.failed = title, [], []

    def check(self, name: str, got: torch.Tensor, want: torch.Tensor, rtol: float, atol: float) -> None:
        got, want = got.detach().float().cpu(), want.detach().float().cpu()
        if got.shape != want.shape:
            self.rows.append(f"{name:28s} SHAPE {tuple(got.shape)} vs {tuple(want.shape)}")
            self.failed.append(name)
            return
        err = (got - want).abs()
        tol = atol + rtol * want.abs()
        ok = bool((err <= tol).all()) and bool(torch.isfinite(got).all())
        rel = float((err / (want.abs() + atol)).max()) if err.numel() else 0.0
        where = ""
        if not ok and err.numel():
            where = f" first-bad-index {tuple(int(i) for i in torch.nonzero(~(err <= tol))[0])}"
        self.rows.append(f"{name:28s} max_abs {float(err.max()) if err.numel() else 0:.3e} max_rel {rel:.3e} scale {float(want.abs().max()):.3e} {'ok' if ok else 'FAIL'}{where}")
        if not ok:
            self.failed.append(name)

    def finish(self) -> None:
        print(f"\n== {self.title}\n" + "\n".join(self.rows))
        assert not self.failed, f"{self.title}: mismatches in {self.failed}"


# ---- tiny encoder/decoder modules with the same parameter names as tests/golden/make_golden.py's stand-ins -------
class LinEncoder(torch.nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.lin = torch.nn.Linear(32 * 32, 64)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.lin(x.flatten(start_dim=-3))


class LinDecoder(torch.nn.Module):
    def __init__(self, feature: int) -> None:
        super().__init__()
        self.lin = torch.nn.Linear(feature, 32 * 32)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.tanh(self.lin(x)).reshape(*x.shape[:-1], 1, 32, 32)


class RandQueue:
    """Context manager: `torch.rand` pops pre-recorded uniforms (golden noise) in call order."""

    def __init__(self, values: list[torch.Tensor]) -> None:
        self.values = list(values)

    def __enter__(self):
        self._orig = torch.rand

        def fake(*size, device=None, dtype=None, generator=None, **_kw):
            v = self.values.pop(0)
            shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
            assert tuple(v.shape) == shape, (tuple(v.shape), shape)
            return v.to(device=device) if device is not None else v.clone()

        torch.rand = fake
        return self

    def __exit__(self, *exc):
        torch.rand = self._orig
        return False


def build_mrssm_model(D: int = 32):
    """Product MoPoE_MRSSM at the default.yaml sizes (D = 32; D = 512 is BASELINE.json cfg3) with LinEncoder/LinDecoder
    (golden-compatible state_dict)."""
    from multimodal_mtrssm_b200.mlp import MLP
    from multimodal_mtrssm_b200.mopoe_mrssm import MoPoE_MRSSM
    from multimodal_mtrssm_b200.networks import Representation, Transition

    rep = dict(deterministic_size=D, hidden_size=D, obs_embed_size=64, distribution_config=[4, 4], activation_name="ELU")
    return MoPoE_MRSSM(
        audio_representation=Representation(**rep), vision_representation=Representation(**rep),
        transition=Transition(deterministic_size=D, hidden_size=D, action_size=6, distribution_config=[4, 4], activation_name="ELU"),
        audio_encoder=LinEncoder(), vision_encoder=LinEncoder(), audio_decoder=LinDecoder(D + 16), vision_decoder=LinDecoder(D + 16),
        init_proj=MLP(in_features=64, out_features=D, num_cells=200, depth=1), kl_coeff=1, use_kl_balancing=True,
    )


def build_mtrssm_model():
    from multimodal_mtrssm_b200.distribution import MultiOneHotFactory
    from multimodal_mtrssm_b200.mlp import MLP
    from multimodal_mtrssm_b200.mopoe_mmtrssm import MoPoE_MMTRSSM
    from multimodal_mtrssm_b200.networks import Representation

    rep = dict(deterministic_size=32, hidden_size=32, obs_embed_size=64, distribution_config=[4, 4], activation_name="ELU")
    head = lambda i: MLP(in_features=i, out_features=16, num_cells=32, depth=1, activation_class=torch.nn.ELU)  # noqa: E731
    return MoPoE_MMTRSSM(
        audio_representation=Representation(**rep), vision_representation=Representation(**rep),
        audio_encoder=LinEncoder(), vision_encoder=LinEncoder(), audio_decoder=LinDecoder(96), vision_decoder=LinDecoder(96),
        init_proj=MLP(in_features=64, out_features=64, num_cells=200, depth=1), kl_coeff=1, use_kl_balancing=True,
        action_size=6, hd_dim=32, hs_dim=16, ld_dim=32, ls_dim=16, l_tau=2.0, h_tau=4.0,
        l_prior=head(32), l_posterior=head(96), h_prior=head(32), h_posterior=head(64),
        l_dist=MultiOneHotFactory(class_size=4, category_size=4), h_dist=MultiOneHotFactory(class_size=2, category_size=8), w_kl_h=1.0,
    )
