"""`MTState`, `stack_mtstates`, `cat_mtstates` -- mirror of the reference's `models/mmtrssm/state.py` (:11-248).

feature order (state.py:51): [deter_h, stoch_h, deter_l, stoch_l].  `hidden_*` is MTRNN.hidden after the update.
As for `State`, a pre-built `feature` may be passed so the fused rollout's single [B,T,96] tensor is shared.
Deviation: `clone()` clones `distribution_h` from `distribution_h` (the reference, state.py:133, clones
`distribution_l` into it -- an upstream slip nothing relies on)."""

from __future__ import annotations

from collections.abc import Generator

import torch
from torch import Tensor

from .distribution import Distribution, cat_distribution, stack_distribution


def _h(t: Tensor, fn) -> Tensor:  # noqa: ANN001  hidden_* may be 1-D in the reference (state.py:77-78)
    return fn(t) if t.dim() > 1 else t


class MTState:
    """Hierarchical latent state with a higher (slow) and a lower (fast) layer (reference: mmtrssm/state.py:11-51)."""

    def __init__(  # noqa: PLR0913
        self,
        deter_h: Tensor,
        deter_l: Tensor,
        distribution_h: Distribution,
        distribution_l: Distribution,
        hidden_h: Tensor,
        hidden_l: Tensor,
        stoch_h: Tensor | None = None,
        stoch_l: Tensor | None = None,
        feature: Tensor | None = None,
    ) -> None:
        self.deter_h, self.deter_l = deter_h, deter_l
        self.distribution_h, self.distribution_l = distribution_h, distribution_l
        self.hidden_h, self.hidden_l = hidden_h, hidden_l
        self.stoch_h = distribution_h.rsample() if stoch_h is None else stoch_h  # state.py:48 (h first)
        self.stoch_l = distribution_l.rsample() if stoch_l is None else stoch_l  # state.py:49
        self.feature = (
            torch.cat([self.deter_h, self.stoch_h, self.deter_l, self.stoch_l], dim=-1) if feature is None else feature
        )

    def __iter__(self) -> Generator["MTState", None, None]:
        for i in range(self.deter_h.shape[0]):
            yield self[i]

    def _map(self, fn, dist_fn) -> "MTState":  # noqa: ANN001
        return type(self)(
            deter_h=fn(self.deter_h), deter_l=fn(self.deter_l),
            distribution_h=dist_fn(self.distribution_h), distribution_l=dist_fn(self.distribution_l),
            hidden_h=_h(self.hidden_h, fn), hidden_l=_h(self.hidden_l, fn),
            stoch_h=fn(self.stoch_h), stoch_l=fn(self.stoch_l),
        )

    def __getitem__(self, loc) -> "MTState":  # noqa: ANN001
        return self._map(lambda t: t[loc], lambda d: d[loc])

    def to(self, device) -> "MTState":  # noqa: ANN001
        return type(self)(
            deter_h=self.deter_h.to(device), deter_l=self.deter_l.to(device),
            distribution_h=self.distribution_h.to(device), distribution_l=self.distribution_l.to(device),
            hidden_h=self.hidden_h.to(device), hidden_l=self.hidden_l.to(device),
            stoch_h=self.stoch_h.to(device), stoch_l=self.stoch_l.to(device),
        )

    def detach(self) -> "MTState":
        return type(self)(
            deter_h=self.deter_h.detach(), deter_l=self.deter_l.detach(),
            distribution_h=self.distribution_h.detach(), distribution_l=self.distribution_l.detach(),
            hidden_h=self.hidden_h.detach(), hidden_l=self.hidden_l.detach(),
            stoch_h=self.stoch_h.detach(), stoch_l=self.stoch_l.detach(),
        )

    def clone(self) -> "MTState":
        return type(self)(
            deter_h=self.deter_h.clone(), deter_l=self.deter_l.clone(),
            distribution_h=self.distribution_h.clone(), distribution_l=self.distribution_l.clone(),
            hidden_h=self.hidden_h.clone(), hidden_l=self.hidden_l.clone(),
            stoch_h=self.stoch_h.clone(), stoch_l=self.stoch_l.clone(),
        )

    def squeeze(self, dim: int) -> "MTState":
        return self._map(lambda t: t.squeeze(dim), lambda d: d.squeeze(dim))

    def unsqueeze(self, dim: int) -> "MTState":
        return self._map(lambda t: t.unsqueeze(dim), lambda d: d.unsqueeze(dim))


def stack_mtstates(states: list[MTState], dim: int) -> MTState:
    """reference: mmtrssm/state.py:184-215"""
    st = lambda name: torch.stack([getattr(s, name) for s in states], dim=dim)  # noqa: E731
    return MTState(
        deter_h=st("deter_h"), deter_l=st("deter_l"), stoch_h=st("stoch_h"), stoch_l=st("stoch_l"),
        distribution_h=stack_distribution([s.distribution_h for s in states], dim),
        distribution_l=stack_distribution([s.distribution_l for s in states], dim),
        hidden_h=st("hidden_h") if states[0].hidden_h.dim() > 1 else states[0].hidden_h,
        hidden_l=st("hidden_l") if states[0].hidden_l.dim() > 1 else states[0].hidden_l,
    )


def cat_mtstates(states: list[MTState], dim: int) -> MTState:
    """reference: mmtrssm/state.py:218-248 (hidden_* taken from the LAST element, :237-238)"""
    ct = lambda name: torch.cat([getattr(s, name) for s in states], dim=dim)  # noqa: E731
    return MTState(
        deter_h=ct("deter_h"), deter_l=ct("deter_l"), stoch_h=ct("stoch_h"), stoch_l=ct("stoch_l"),
        distribution_h=cat_distribution([s.distribution_h for s in states], dim),
        distribution_l=cat_distribution([s.distribution_l for s in states], dim),
        hidden_h=states[-1].hidden_h, hidden_l=states[-1].hidden_l,
    )
