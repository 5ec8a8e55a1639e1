"""`State`, `stack_states`, `cat_states` -- mirror of the reference's `models/state.py` (:11-152).

Same attributes (`deter`, `distribution`, `stoch`, `feature`) and container algebra.  One addition: the constructor
accepts a pre-built `feature`, so the fused rollout can hand out ONE [B,T,D+S] tensor with `deter` / `stoch` as views
of it instead of a T-way stack followed by a concat (state.py:18,132-134)."""

from __future__ import annotations

from collections.abc import Generator

import torch
from torch import Tensor

from .distribution import Distribution, cat_distribution, stack_distribution


class State:
    """Latent state with deterministic and stochastic parts (reference: models/state.py:11-18)."""

    def __init__(self, deter: Tensor, distribution: Distribution, stoch: Tensor | None = None, feature: Tensor | None = None) -> None:
        self.deter = deter
        self.distribution = distribution
        self.stoch = distribution.rsample() if stoch is None else stoch  # state.py:17: samples in the constructor
        self.feature = torch.cat([self.deter, self.stoch], dim=-1) if feature is None else feature  # state.py:18

    def __iter__(self) -> Generator["State", None, None]:
        for i in range(self.deter.shape[0]):
            yield self[i]

    def __getitem__(self, loc) -> "State":  # noqa: ANN001
        return type(self)(deter=self.deter[loc], stoch=self.stoch[loc], distribution=self.distribution[loc])

    def to(self, device) -> "State":  # noqa: ANN001
        return type(self)(deter=self.deter.to(device), stoch=self.stoch.to(device), distribution=self.distribution.to(device))

    def detach(self) -> "State":
        return type(self)(deter=self.deter.detach(), stoch=self.stoch.detach(), distribution=self.distribution.detach())

    def clone(self) -> "State":
        return type(self)(deter=self.deter.clone(), stoch=self.stoch.clone(), distribution=self.distribution.clone())

    def squeeze(self, dim: int) -> "State":
        return type(self)(deter=self.deter.squeeze(dim), stoch=self.stoch.squeeze(dim), distribution=self.distribution.squeeze(dim))

    def unsqueeze(self, dim: int) -> "State":
        return type(self)(
            deter=self.deter.unsqueeze(dim), stoch=self.stoch.unsqueeze(dim), distribution=self.distribution.unsqueeze(dim)
        )


def stack_states(states: list[State], dim: int) -> State:
    """reference: models/state.py:121-135"""
    return State(
        deter=torch.stack([s.deter for s in states], dim=dim),
        stoch=torch.stack([s.stoch for s in states], dim=dim),
        distribution=stack_distribution([s.distribution for s in states], dim),
    )


def cat_states(states: list[State], dim: int) -> State:
    """reference: models/state.py:138-152"""
    return State(
        deter=torch.cat([s.deter for s in states], dim=dim),
        stoch=torch.cat([s.stoch for s in states], dim=dim),
        distribution=cat_distribution([s.distribution for s in states], dim),
    )
