"""`BaseRSSM` -- mirror of the reference's `models/core.py` (:13-266).

Same constructor, abstract hooks, `initial_state`, `rollout_representation`, `rollout_transition`, `shared_step`,
`training_step`, `validation_step` and logged keys.  It is a `lightning.LightningModule` when lightning is
installed, otherwise an `nn.Module` with the two members the methods need (`device`, `log_dict`), so the same class
is driven by LightningCLI there and by a plain loop here."""

from __future__ import annotations

from abc import abstractmethod

import torch
from torch import Tensor, nn

from .distribution import kl_divergence
from .networks import Representation, Transition
from .state import State

try:  # pragma: no cover - lightning is absent in the build image
    from lightning import LightningModule as _Base
except ImportError:

    class _Base(nn.Module):  # type: ignore[no-redef]
        """Minimal LightningModule surface used by the RSSM classes."""

        @property
        def device(self) -> torch.device:
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log_dict(self, *_args, **_kwargs) -> None:  # noqa: ANN002, ANN003
            return None


class BaseRSSM(_Base):
    """Base RSSM (reference: models/core.py:13-31)."""

    #: None = follow autocast (bf16 tensor-core path when `torch.is_autocast_enabled()`, else the fp32-parity path);
    #: "fp32" / "bf16" force one.
    rollout_precision: str | None = None

    def __init__(
        self,
        *,
        representation: Representation,
        transition: Transition,
        init_proj: nn.Module,
        kl_coeff: float,
        use_kl_balancing: bool,
    ) -> None:
        super().__init__()
        self.representation = representation
        self.transition = transition
        self.init_proj = init_proj
        self.kl_coeff = kl_coeff
        self.use_kl_balancing = use_kl_balancing

    # ---- hooks the variants implement (core.py:33-119) --------------------------------------------------------
    @abstractmethod
    def encode_observation(self, observation): ...  # noqa: ANN001, ANN201

    @abstractmethod
    def decode_state(self, state): ...  # noqa: ANN001, ANN201

    @abstractmethod
    def compute_reconstruction_loss(self, reconstructions, targets): ...  # noqa: ANN001, ANN201

    @abstractmethod
    def get_observations_from_batch(self, batch): ...  # noqa: ANN001, ANN201

    @abstractmethod
    def get_initial_observation(self, observations): ...  # noqa: ANN001, ANN201

    @abstractmethod
    def get_targets_from_batch(self, batch): ...  # noqa: ANN001, ANN201

    # ---- precision policy -----------------------------------------------------------------------------------------
    def _precision(self) -> int:
        from . import _lib

        # the bf16 tensor-core path of this model family (MoPoE-MMTRSSM overrides it with the fused-backward policy)
        bf16 = getattr(self, "_BF16_POLICY", _lib.PRECISION_BF16)
        if self.rollout_precision is not None:
            return {"fp32": _lib.PRECISION_FP32, "bf16": bf16, "bf16_two_kernel": _lib.PRECISION_BF16}[self.rollout_precision]
        return bf16 if torch.is_autocast_enabled() else _lib.PRECISION_FP32

    # ---- reference API ----------------------------------------------------------------------------------------------
    def initial_state(self, observation) -> State:  # noqa: ANN001
        """core.py:121-135: encode -> init_proj -> prior head -> factory -> State (samples)."""
        deter = self.init_proj(self.encode_observation(observation))
        logits = self.transition.rnn_to_prior_projector(deter)
        return State(deter=deter, distribution=self.representation.distribution_factory(logits)).to(self.device)

    def rollout_representation(self, *, actions: Tensor, observations, prev_state: State) -> tuple[State, State]:  # noqa: ANN001
        """core.py:137-168, the unimodal rollout (both shipped models override it).  The T loop is ONE fused kernel
        (`rollout_ops.mrssm_rollout(..., unimodal=True)`: Transition.forward + ONE Representation.forward per step, no fusion).
        There is no per-step fallback: CPU tensors, sizes outside the built family (include/rssm_rollout.h) or heads that are
        not Linear-ELU-Linear raise RuntimeError."""
        obs_embed = self.encode_observation(observations)
        return self._rollout_representation_fused(actions, obs_embed, prev_state)

    def _unimodal_weights(self) -> list[Tensor]:
        """Parameters in C-ABI order (`params.MR_STATE_KEYS`); the single head fills both posterior slots (the second is unused
        by the unimodal kernels and gets zero gradients)."""
        from .mopoe_mrssm import mlp_params

        tr = self.transition
        for name in ("action_state_projector", "rnn_cell", "rnn_to_prior_projector"):
            if not hasattr(tr, name):
                raise RuntimeError(f"the fused rollout needs a `Transition` with `{name}` (networks.py:126-149), got {type(tr).__name__}")
        head = mlp_params(self.representation.rnn_to_post_projector, "representation.rnn_to_post_projector")
        return [
            *mlp_params(tr.action_state_projector, "transition.action_state_projector"),
            tr.rnn_cell.weight_ih, tr.rnn_cell.weight_hh, tr.rnn_cell.bias_ih, tr.rnn_cell.bias_hh,
            *mlp_params(tr.rnn_to_prior_projector, "transition.rnn_to_prior_projector"), *head, *head,
        ]

    def _rollout_representation_fused(self, actions: Tensor, obs_embed: Tensor, prev_state: State) -> tuple[State, State]:
        from . import rollout_ops
        from .distribution import Distribution, FusedKL
        from .mopoe_mrssm import flat_stoch

        f = self.representation.distribution_factory
        B, T = obs_embed.shape[:2]
        C, dev = int(f.category_size), obs_embed.device
        out = rollout_ops.mrssm_rollout(
            self._unimodal_weights(), actions=actions, embed_a=obs_embed, embed_v=obs_embed, h0=prev_state.deter,
            z0=flat_stoch(prev_state.stoch), u_post=torch.rand(B, T, C, device=dev), u_prior=torch.rand(B, T, C, device=dev),
            class_size=int(f.class_size), precision=self._precision(), use_kl_balancing=bool(self.use_kl_balancing), unimodal=True,
        )
        feature = out["feature"]
        D = prev_state.deter.shape[-1]
        link = FusedKL(kl=out["kl"], use_balancing=bool(self.use_kl_balancing), token=object())
        posterior = State(deter=feature[..., :D], stoch=feature[..., D:], feature=feature,
                          distribution=Distribution(out["post_probs"], _fused=link, _role="post"))
        prior = State(deter=feature[..., :D], stoch=out["prior_stoch"], distribution=Distribution(out["prior_probs"], _fused=link, _role="prior"))
        return posterior, prior

    def rollout_transition(self, *, actions: Tensor, prev_state: State) -> State:
        """core.py:170-185: imagination, the prior's own sample is fed back.  ONE fused forward-only kernel
        (`rollout_ops.mrssm_imagine`; the reference calls this under no_grad); no per-step fallback."""
        from . import rollout_ops
        from .distribution import Distribution
        from .mopoe_mrssm import flat_stoch

        if torch.is_grad_enabled() and (actions.requires_grad or prev_state.deter.requires_grad):
            msg = "the fused rollout_transition is forward-only; call it under torch.no_grad() (as the reference's callbacks do)"
            raise RuntimeError(msg)
        B, T = actions.shape[:2]
        fac = self.transition.distribution_factory
        u = torch.rand(B, T, int(fac.category_size), device=actions.device)
        out = rollout_ops.mrssm_imagine(
            [w.detach() for w in self._unimodal_weights()], actions=actions, h0=prev_state.deter, z0=flat_stoch(prev_state.stoch), u=u,
            class_size=int(fac.class_size), precision=self._precision(),
        )
        feature = out["feature"]
        D = prev_state.deter.shape[-1]
        return State(deter=feature[..., :D], stoch=feature[..., D:], feature=feature, distribution=Distribution(out["probs"]))

    def shared_step(self, batch: tuple[Tensor, ...]) -> dict[str, Tensor]:
        """core.py:187-221: recon + kl_coeff * KL(posterior || prior)."""
        observations = self.get_observations_from_batch(batch)
        posterior, prior = self.rollout_representation(
            actions=batch[0],  # the action is always first (core.py:197)
            observations=observations,
            prev_state=self.initial_state(self.get_initial_observation(observations)),
        )
        loss_dict = self.compute_reconstruction_loss(self.decode_state(posterior), self.get_targets_from_batch(batch))
        kl = kl_divergence(
            q=posterior.distribution.independent(1), p=prior.distribution.independent(1), use_balancing=self.use_kl_balancing
        ).mul(self.kl_coeff)
        loss_dict["kl"] = kl
        loss_dict["loss"] = loss_dict["recon"] + kl
        return loss_dict

    def _step(self, batch: tuple[Tensor, ...], stage: str) -> dict[str, Tensor]:
        loss_dict = self.shared_step(batch)
        out = {f"{stage}/{k}": v for k, v in loss_dict.items()}
        if stage == "train":
            out = {"loss": loss_dict["loss"], **out}  # Lightning's automatic optimisation needs "loss" (core.py:234-236)
        self.log_dict(out, prog_bar=True, sync_dist=True, on_step=False, on_epoch=True)
        return out

    def training_step(self, batch: tuple[Tensor, ...], _: int) -> dict[str, Tensor]:
        """core.py:223-244"""
        return self._step(batch, "train")

    def validation_step(self, batch: tuple[Tensor, ...], _batch_index: int) -> dict[str, Tensor]:
        """core.py:246-266"""
        return self._step(batch, "val")
