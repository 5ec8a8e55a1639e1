"""Stand-in for the `distribution_extension` names the hot path uses (absent third-party package, pinned at
nomutin/distribution-extension@e150621, `uv.lock:744-746`): `MultiOneHotFactory`, `Distribution`,
`kl_divergence`, `stack_distribution`, `cat_distribution`.  Semantics = assumptions A1..A5 of SURVEY.md §8(c).

When the real package is importable, `compat.install()` leaves it alone and these classes are only used for the
distributions the fused rollout returns (they duck-type the same interface).

Fused KL: distributions produced by one fused rollout carry a link to the kernel's per-(b,t) KL tensor; calling
`kl_divergence(q=post.independent(1), p=prior.independent(1), use_balancing=...)` on the untouched pair returns its
mean (gradient routed inside the backward kernel) instead of re-deriving it from the probabilities.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributions as td
from torch import Tensor, nn

KL_BALANCE_ALPHA = 0.8  # A5


@dataclass
class FusedKL:
    """Per-(b,t) KL(post || prior) summed over groups, computed by the rollout kernel."""

    kl: Tensor  # [B, T]
    use_balancing: bool  # the gradient routing the kernel will apply
    token: object


class Distribution:
    """Multi-categorical (C groups x K classes) parameterised by probabilities [..., C, K] (A1, A3)."""

    def __init__(self, probs: Tensor, *, _fused: FusedKL | None = None, _role: str = "") -> None:
        self.probs = probs
        self._fused, self._role = _fused, _role

    @property
    def parameters(self) -> dict[str, Tensor]:
        return {"probs": self.probs}

    # A4
    def independent(self, dim: int) -> td.Independent:
        # OneHotCategorical builds its inner Categorical WITHOUT forwarding validate_args, and that validation reads a device
        # flag on the host (`torch._is_all_true`): a sync per call and illegal under CUDA-graph capture -> switch the default off
        # around the construction (the probabilities come from the kernels' / factory's softmax)
        default = td.Distribution._validate_args  # noqa: SLF001
        td.Distribution.set_default_validate_args(False)
        try:
            ind = td.Independent(td.OneHotCategoricalStraightThrough(probs=self.probs, validate_args=False), dim, validate_args=False)
        finally:
            td.Distribution.set_default_validate_args(default)
        ind._rssm_fused, ind._rssm_role = self._fused, self._role  # type: ignore[attr-defined]
        return ind

    # A2: straight-through one-hot, flattened to [..., C*K]; inverse-CDF draw from the global torch RNG
    def rsample(self) -> Tensor:
        probs = self.probs
        u = torch.rand(probs.shape[:-1], device=probs.device, dtype=probs.dtype)
        idx = (probs.detach().cumsum(-1) <= u.unsqueeze(-1)).sum(-1).clamp(max=probs.shape[-1] - 1)
        onehot = torch.nn.functional.one_hot(idx, probs.shape[-1]).to(probs.dtype)
        return (onehot + probs - probs.detach()).flatten(start_dim=-2)

    def sample(self) -> Tensor:
        return self.rsample().detach()

    # A3: container algebra on the parameter tensor along batch dims (any of these drops the fused-KL link)
    def __getitem__(self, loc) -> "Distribution":  # noqa: ANN001
        return type(self)(self.probs[loc])

    def to(self, device) -> "Distribution":  # noqa: ANN001
        moved = self.probs.to(device)
        return self if moved is self.probs else type(self)(moved)

    def detach(self) -> "Distribution":
        return type(self)(self.probs.detach())

    def clone(self) -> "Distribution":
        return type(self)(self.probs.clone())

    def squeeze(self, dim: int) -> "Distribution":
        return type(self)(self.probs.squeeze(dim))

    def unsqueeze(self, dim: int) -> "Distribution":
        return type(self)(self.probs.unsqueeze(dim))


MultiOneHot = Distribution


class MultiOneHotFactory(nn.Module):
    """A1: logits [..., class_size*category_size] -> [..., category_size, class_size], softmax over classes."""

    def __init__(self, class_size: int, category_size: int) -> None:
        super().__init__()
        self.class_size = class_size
        self.category_size = category_size

    def forward(self, logits: Tensor) -> Distribution:
        shaped = logits.reshape(*logits.shape[:-1], self.category_size, self.class_size)
        return Distribution(torch.softmax(shaped, dim=-1))


def _probs_of(d) -> Tensor:  # noqa: ANN001
    base = d.base_dist if isinstance(d, td.Independent) else d
    return base.probs


def kl_divergence(*, q, p, use_balancing: bool) -> Tensor:  # noqa: ANN001
    """A5: mean over batch dims of KL(q || p) (summed over groups); balanced = a*KL(sg q||p) + (1-a)*KL(q||sg p)."""
    fq, fp = getattr(q, "_rssm_fused", None), getattr(p, "_rssm_fused", None)
    if (fq is not None and fp is not None and fq.token is fp.token and q._rssm_role == "post" and p._rssm_role == "prior"
            and fq.use_balancing == bool(use_balancing)):
        return fq.kl.mean()
    qp, pp = _probs_of(q), _probs_of(p)

    def kl(a: Tensor, b: Tensor) -> Tensor:
        eps = torch.finfo(a.dtype).eps
        return (a * (a.clamp(eps, 1 - eps).log() - b.clamp(eps, 1 - eps).log())).sum(-1).sum(-1).mean()

    if not use_balancing:
        return kl(qp, pp)
    return KL_BALANCE_ALPHA * kl(qp.detach(), pp) + (1 - KL_BALANCE_ALPHA) * kl(qp, pp.detach())


def stack_distribution(dists: list[Distribution], dim: int) -> Distribution:
    return Distribution(torch.stack([d.probs for d in dists], dim=dim))


def cat_distribution(dists: list[Distribution], dim: int) -> Distribution:
    return Distribution(torch.cat([d.probs for d in dists], dim=dim))
