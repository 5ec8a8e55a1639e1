"""Synthetic workloads for the rollout: default-init weights (seed 42), data (seed 1234),
categorical noise (seed 4321).  Used by bench.py, the examples and the tests; everything is generated on CPU
with fixed generators so that every device / rank sees identical numbers."""

from __future__ import annotations

import torch
from torch import Tensor

from .params import MR_STATE_KEYS, MT_STATE_KEYS

MR_SHAPES = {
    "transition.action_state_projector.0": (32, 22), "transition.action_state_projector.2": (32, 32),
    "transition.rnn_to_prior_projector.0": (32, 32), "transition.rnn_to_prior_projector.2": (16, 32),
    "audio_representation.rnn_to_post_projector.0": (32, 96), "audio_representation.rnn_to_post_projector.2": (16, 32),
    "vision_representation.rnn_to_post_projector.0": (32, 96), "vision_representation.rnn_to_post_projector.2": (16, 32),
}
MT_SHAPES = {
    "l_rnn._d2h": (32, 32), "l_rnn._input2h": (32, 38), "h_rnn._d2h": (32, 32), "h_rnn._input2h": (32, 16),
    "l_prior.0": (32, 32), "l_prior.2": (16, 32), "h_prior.0": (32, 32), "h_prior.2": (16, 32),
    "h_posterior.0": (32, 64), "h_posterior.2": (16, 32),
    "audio_representation.rnn_to_post_projector.0": (32, 96), "audio_representation.rnn_to_post_projector.2": (16, 32),
    "vision_representation.rnn_to_post_projector.0": (32, 96), "vision_representation.rnn_to_post_projector.2": (16, 32),
}


def _linear_init(shapes: dict, g: torch.Generator, gain: float) -> dict[str, Tensor]:
    out = {}
    for name, (o, i) in shapes.items():
        bound = gain / i**0.5  # nn.Linear default: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
        out[f"{name}.weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        out[f"{name}.bias"] = (torch.rand(o, generator=g) * 2 - 1) * bound
    return out


def mrssm_shapes(D: int, A: int = 6) -> dict:
    """Linear shapes of MoPoE-MRSSM for deterministic_size = hidden_size = D (D = 32: default.yaml; D = 512: BASELINE cfg3)."""
    t, a, v = "transition.", "audio_representation.rnn_to_post_projector.", "vision_representation.rnn_to_post_projector."
    return {
        t + "action_state_projector.0": (D, A + 16), t + "action_state_projector.2": (D, D),
        t + "rnn_to_prior_projector.0": (D, D), t + "rnn_to_prior_projector.2": (16, D),
        a + "0": (D, D + 64), a + "2": (16, D), v + "0": (D, D + 64), v + "2": (16, D),
    }


def mrssm_params(seed: int = 42, gain: float = 1.0, D: int = 32) -> dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    p = _linear_init(MR_SHAPES if D == 32 else mrssm_shapes(D), g, gain)
    b = gain / D**0.5  # nn.GRUCell default: U(-1/sqrt(hidden), 1/sqrt(hidden))
    for k, s in (("weight_ih", (3 * D, D)), ("weight_hh", (3 * D, D)), ("bias_ih", (3 * D,)), ("bias_hh", (3 * D,))):
        p[f"transition.rnn_cell.{k}"] = (torch.rand(*s, generator=g) * 2 - 1) * b
    assert set(p) == set(MR_STATE_KEYS)
    return p


def mtrssm_params(seed: int = 42, gain: float = 1.0) -> dict[str, Tensor]:
    p = _linear_init(MT_SHAPES, torch.Generator().manual_seed(seed), gain)
    assert set(p) == set(MT_STATE_KEYS)
    return p


def actions(B: int, T: int, g: torch.Generator, A: int = 6) -> Tensor:
    """One-hot speaker id, constant over t (scripts/convert_audio_mnist_data.py:35) + N(0, 0.1^2) (transform.py:55-72)."""
    speaker = torch.randint(0, A, (B,), generator=g)
    act = torch.nn.functional.one_hot(speaker, A).float()[:, None, :].expand(B, T, A)
    return (act + 0.1 * torch.randn(B, T, A, generator=g)).contiguous()


def onehot(B: int, C: int, K: int, g: torch.Generator) -> Tensor:
    return torch.nn.functional.one_hot(torch.randint(0, K, (B, C), generator=g), K).float().flatten(1)


def mtrssm_batch(B: int, T: int, seed: int = 1234, noise_seed: int = 4321, prior_noise: bool = False) -> dict[str, Tensor]:
    """Rollout-only MMTRSSM inputs at the default.yaml sizes (l_dist 4x4, h_dist 8 groups x 2 classes).  `prior_noise` adds the
    uniforms of the prior MTState's own draws (mmtrssm/state.py:48-49), which `MoPoE_MMTRSSM.rollout_representation` always makes."""
    g, n = torch.Generator().manual_seed(seed), torch.Generator().manual_seed(noise_seed)
    d_h, d_l = torch.randn(B, 32, generator=g), torch.randn(B, 32, generator=g)
    out = {
        "actions": actions(B, T, g), "embed_a": torch.randn(B, T, 64, generator=g), "embed_v": torch.randn(B, T, 64, generator=g),
        "deter_h0": d_h, "deter_l0": d_l, "hidden_h0": d_h.clone(), "hidden_l0": d_l.clone(),
        "stoch_h0": onehot(B, 8, 2, g), "stoch_l0": onehot(B, 4, 4, g),
        "u_post_l": torch.rand(B, T, 4, generator=n), "u_post_h": torch.rand(B, T, 8, generator=n),
    }
    if prior_noise:
        out["u_prior_l"], out["u_prior_h"] = torch.rand(B, T, 4, generator=n), torch.rand(B, T, 8, generator=n)
    return out


def mrssm_batch(B: int, T: int, seed: int = 1234, noise_seed: int = 4321, D: int = 32) -> dict[str, Tensor]:
    g, n = torch.Generator().manual_seed(seed), torch.Generator().manual_seed(noise_seed)
    return {
        "actions": actions(B, T, g), "embed_a": torch.randn(B, T, 64, generator=g), "embed_v": torch.randn(B, T, 64, generator=g),
        "h0": torch.randn(B, D, generator=g), "z0": onehot(B, 4, 4, g), "u_post": torch.rand(B, T, 4, generator=n),
    }
