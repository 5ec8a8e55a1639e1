"""`Representation`, `Transition` -- mirror of the reference's `models/networks.py` (:18-173).

Same constructor signatures, sub-module names (hence `state_dict` keys) and single-step `forward` semantics.  The
single-step forwards stay plain PyTorch (they are the per-step API callers may still use); the T-step rollouts of the
model classes do not call them -- they hand these modules' parameters to the fused CUDA rollout."""

from __future__ import annotations

import torch
from torch import Tensor, nn

from .distribution import MultiOneHotFactory
from .mlp import MLP
from .state import State


def _distribution_config(cfg: tuple[int, int] | list[int]) -> tuple[int, int]:
    """(class_size, category_size); YAML hands a list (networks.py:46-54)."""
    if isinstance(cfg, list):
        if len(cfg) != 2:  # noqa: PLR2004
            msg = f"distribution_config must have 2 elements, got {len(cfg)}"
            raise ValueError(msg)
        return cfg[0], cfg[1]
    return cfg


class Representation(nn.Module):
    """Posterior head q(z_t | h_t, e_t) (reference: networks.py:18-84)."""

    def __init__(
        self,
        *,
        deterministic_size: int,
        hidden_size: int,
        obs_embed_size: int,
        distribution_config: tuple[int, int] | list[int],
        activation_name: str = "ReLU",
    ) -> None:
        super().__init__()
        class_size, category_size = _distribution_config(distribution_config)
        self.activation_name = activation_name
        self.rnn_to_post_projector = MLP(
            in_features=obs_embed_size + deterministic_size,
            out_features=class_size * category_size,
            num_cells=hidden_size,
            depth=1,
            activation_class=getattr(nn, activation_name),
            activate_last_layer=False,
        )
        self.distribution_factory = MultiOneHotFactory(class_size=class_size, category_size=category_size)

    def forward(self, obs_embed: Tensor, prior_state: State) -> State:
        logits = self.rnn_to_post_projector(torch.cat([prior_state.deter, obs_embed], -1))
        return State(deter=prior_state.deter, distribution=self.distribution_factory(logits))


class Transition(nn.Module):
    """GRU transition + prior head p(z_t | h_t) (reference: networks.py:87-173)."""

    def __init__(
        self,
        *,
        deterministic_size: int,
        hidden_size: int,
        action_size: int,
        distribution_config: tuple[int, int] | list[int],
        activation_name: str,
    ) -> None:
        super().__init__()
        class_size, category_size = _distribution_config(distribution_config)
        self.activation_name = activation_name
        act = getattr(nn, activation_name)
        self.rnn_cell = nn.GRUCell(input_size=hidden_size, hidden_size=deterministic_size)
        self.action_state_projector = MLP(
            in_features=action_size + class_size * category_size, out_features=hidden_size, num_cells=hidden_size, depth=1,
            activation_class=act, activate_last_layer=False,
        )
        self.rnn_to_prior_projector = MLP(
            in_features=deterministic_size, out_features=class_size * category_size, num_cells=hidden_size, depth=1,
            activation_class=act, activate_last_layer=False,
        )
        self.distribution_factory = MultiOneHotFactory(class_size=class_size, category_size=category_size)

    def forward(self, action: Tensor, prev_state: State) -> State:
        stoch = prev_state.stoch.flatten(start_dim=1) if prev_state.stoch.dim() >= 3 else prev_state.stoch  # noqa: PLR2004
        deter = self.rnn_cell(self.action_state_projector(torch.cat([action, stoch], dim=-1)), prev_state.deter)
        return State(deter=deter, distribution=self.distribution_factory(self.rnn_to_prior_projector(deter)))
