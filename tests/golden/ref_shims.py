"""Stand-ins for the reference's ABSENT third-party dependencies (test infrastructure only).

The reference (`/root/reference/src/multimodal_rssm`) cannot be imported in this image because
`distribution_extension`, `torchrl`, `lightning` and `cnn` are not installed.
This module registers minimal stand-ins for exactly the names the hot path imports, so that the
reference's OWN rollout code (models/core.py, networks.py, state.py, mopoe_*/core.py) executes
unmodified and can generate golden vectors (`make_golden.py`).

What is restated here is third-party behaviour, not reference code:

* `distribution_extension` 1.0.7 (nomutin/distribution-extension @ e150621, `uv.lock:744-746`):
  `MultiOneHotFactory`, `Distribution`, `kl_divergence`, `utils.stack_distribution`,
  `utils.cat_distribution` -- assumptions A1..A5 of SURVEY.md §8(c).
* `torchrl.modules.MLP` 0.10.1 -- assumption A6 (`Sequential(Linear, act, Linear)`, default Tanh).
* `lightning.LightningModule` -- reduced to `nn.Module` + `device` + no-op `log_dict`.

Sampling: the reference draws with the global torch RNG inside `State.__init__`
(`models/state.py:17`).  To make the noise explicit, `rsample()` here pops uniforms from
`NOISE` (a FIFO the golden script fills and records) and draws by inverse CDF:
`idx = min(K-1, #{k : cdf_k <= u})`.

Nothing under `multimodal_mtrssm_b200/` may import this file.
"""

from __future__ import annotations

import importlib
import sys
import types
from collections import deque

import torch
import torch.distributions as td
from torch import Tensor, nn

REFERENCE_SRC = "/root/reference/src"


class NoiseQueue:
    """FIFO of uniforms consumed by `MultiOneHot.rsample`; records (tag, u) of every draw."""

    def __init__(self) -> None:
        self.generator: torch.Generator | None = None
        self.forced: deque[Tensor] = deque()
        self.log: list[Tensor] = []

    def reset(self, seed: int) -> None:
        self.generator = torch.Generator().manual_seed(seed)
        self.forced.clear()
        self.log = []

    def draw(self, shape: torch.Size) -> Tensor:
        if self.forced:
            u = self.forced.popleft()
            assert u.shape == shape, (u.shape, shape)
        else:
            assert self.generator is not None, "call NOISE.reset(seed) first"
            u = torch.rand(shape, generator=self.generator)
        self.log.append(u)
        return u


NOISE = NoiseQueue()


class Distribution:
    """A3: container algebra acts on the parameter tensor along batch dims."""

    def __init__(self, probs: Tensor) -> None:
        self.probs = probs  # [..., category(C), class(K)]

    # --- A4 -------------------------------------------------------------------------------
    def independent(self, dim: int) -> td.Independent:
        return td.Independent(td.OneHotCategoricalStraightThrough(probs=self.probs), dim)

    # --- A2 -------------------------------------------------------------------------------
    def rsample(self) -> Tensor:
        probs = self.probs
        u = NOISE.draw(probs.shape[:-1]).to(probs.device, probs.dtype)
        cdf = probs.detach().cumsum(-1)
        idx = (cdf <= u.unsqueeze(-1)).sum(-1).clamp(max=probs.shape[-1] - 1)
        onehot = torch.nn.functional.one_hot(idx, probs.shape[-1]).to(probs.dtype)
        sample = onehot + probs - probs.detach()  # straight-through
        return sample.flatten(start_dim=-2)

    def sample(self) -> Tensor:
        return self.rsample().detach()

    # --- A3 -------------------------------------------------------------------------------
    def __getitem__(self, loc):  # noqa: ANN001
        return type(self)(self.probs[loc])

    def to(self, device):  # noqa: ANN001
        return type(self)(self.probs.to(device))

    def detach(self):
        return type(self)(self.probs.detach())

    def clone(self):
        return type(self)(self.probs.clone())

    def squeeze(self, dim: int):
        return type(self)(self.probs.squeeze(dim))

    def unsqueeze(self, dim: int):
        return type(self)(self.probs.unsqueeze(dim))


class MultiOneHotFactory(nn.Module):
    """A1: logits[..., S] -> [..., category_size, class_size], softmax over the class axis."""

    def __init__(self, class_size: int, category_size: int) -> None:
        super().__init__()
        self.class_size = class_size
        self.category_size = category_size

    def forward(self, logits: Tensor) -> Distribution:
        shaped = logits.reshape(*logits.shape[:-1], self.category_size, self.class_size)
        return Distribution(torch.softmax(shaped, dim=-1))


def kl_divergence(*, q: td.Independent, p: td.Independent, use_balancing: bool) -> Tensor:
    """A5: mean over batch dims; balancing = 0.8*KL(sg q || p) + 0.2*KL(q || sg p)."""
    if not use_balancing:
        return td.kl_divergence(q, p).mean()
    alpha = 0.8

    def sg(d: td.Independent) -> td.Independent:
        return td.Independent(
            td.OneHotCategoricalStraightThrough(probs=d.base_dist.probs.detach()),
            d.reinterpreted_batch_ndims,
        )

    return alpha * td.kl_divergence(sg(q), p).mean() + (1 - alpha) * td.kl_divergence(q, sg(p)).mean()


def stack_distribution(dists: list[Distribution], dim: int) -> Distribution:
    return Distribution(torch.stack([d.probs for d in dists], dim=dim))


def cat_distribution(dists: list[Distribution], dim: int) -> Distribution:
    return Distribution(torch.cat([d.probs for d in dists], dim=dim))


class MLP(nn.Sequential):
    """A6: torchrl.modules.MLP(depth=1) == Sequential(Linear, act, Linear); default act Tanh."""

    def __init__(
        self,
        in_features: int,
        out_features: int,
        num_cells: int,
        depth: int = 1,
        activation_class: type[nn.Module] = nn.Tanh,
        activate_last_layer: bool = False,
    ) -> None:
        assert depth == 1 and not activate_last_layer
        super().__init__(nn.Linear(in_features, num_cells), activation_class(), nn.Linear(num_cells, out_features))


class LightningModule(nn.Module):
    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def log_dict(self, *_args, **_kwargs) -> None:
        return None


def _module(name: str, **attrs) -> types.ModuleType:  # noqa: ANN003
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _package_stub(name: str, path: str) -> None:
    """Register `name` as a package WITHOUT running its __init__ (which imports wandb/gdown/...)."""
    mod = types.ModuleType(name)
    mod.__path__ = [path]  # type: ignore[attr-defined]
    sys.modules[name] = mod


def install() -> None:
    """Register the stand-ins and make the reference's hot-path modules importable."""
    de = _module(
        "distribution_extension",
        Distribution=Distribution,
        MultiOneHotFactory=MultiOneHotFactory,
        kl_divergence=kl_divergence,
    )
    de.utils = _module(
        "distribution_extension.utils",
        stack_distribution=stack_distribution,
        cat_distribution=cat_distribution,
    )
    _module("lightning", LightningModule=LightningModule)
    tr = _module("torchrl")
    tr.modules = _module("torchrl.modules", MLP=MLP)

    base = f"{REFERENCE_SRC}/multimodal_rssm"
    _package_stub("multimodal_rssm", base)
    _package_stub("multimodal_rssm.models", f"{base}/models")
    _package_stub("multimodal_rssm.models.mrssm", f"{base}/models/mrssm")
    _package_stub("multimodal_rssm.models.mrssm.mopoe_mrssm", f"{base}/models/mrssm/mopoe_mrssm")
    _package_stub("multimodal_rssm.models.mmtrssm", f"{base}/models/mmtrssm")
    _package_stub("multimodal_rssm.models.mmtrssm.mopoe_mmtrssm", f"{base}/models/mmtrssm/mopoe_mmtrssm")


def import_reference():
    """Return the reference's own hot-path modules (executed from /root/reference, unmodified)."""
    install()
    names = {
        "core": "multimodal_rssm.models.core",
        "networks": "multimodal_rssm.models.networks",
        "state": "multimodal_rssm.models.state",
        "objective": "multimodal_rssm.models.objective",
        "mopoe_mrssm": "multimodal_rssm.models.mrssm.mopoe_mrssm.core",
        "mtstate": "multimodal_rssm.models.mmtrssm.state",
        "mopoe_mmtrssm": "multimodal_rssm.models.mmtrssm.mopoe_mmtrssm.core",
    }
    return types.SimpleNamespace(**{k: importlib.import_module(v) for k, v in names.items()})
