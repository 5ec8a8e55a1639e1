"""`likelihood` -- mirror of the reference's `models/objective.py:7-23` (same signature, same value)."""

from __future__ import annotations

import math

from torch import Tensor


def likelihood(prediction: Tensor, target: Tensor, event_ndims: int, scale: float = 1.0) -> Tensor:
    """-mean over batch dims of log Independent(Normal(prediction, scale), event_ndims).log_prob(target).

    Closed form of objective.py:21-23: 0.5*||(x-mu)/scale||^2 + n*(log(scale) + 0.5*log(2*pi)) summed over the last
    `event_ndims` dims, then averaged -- one fused reduction instead of building distribution objects.
    """
    dims = tuple(range(-event_ndims, 0))
    n = math.prod(prediction.shape[-event_ndims:])
    sq = ((target - prediction) / scale).square().sum(dims)
    return (0.5 * sq + n * (math.log(scale) + 0.5 * math.log(2 * math.pi))).mean()
