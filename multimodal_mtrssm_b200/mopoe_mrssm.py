"""`MoPoE_MRSSM` -- mirror of the reference's `models/mrssm/mopoe_mrssm/core.py` (:12-355).

Same constructor, module names (`state_dict` keys incl. the `representation.*` alias of `audio_representation.*`),
batch layout and methods.  `rollout_representation` / `rollout_transition` run the fused CUDA rollout
(`rollout_ops.mrssm_rollout` / `mrssm_imagine`) instead of the per-step Python loop (:221-256); encoders and decoders
stay PyTorch modules at the kernel boundary.  There is no eager fallback: CPU tensors or unsupported sizes raise."""

from __future__ import annotations

import torch
from torch import Tensor, nn

from . import rollout_ops
from .core import BaseRSSM
from .distribution import Distribution, FusedKL
from .networks import Representation, Transition
from .objective import likelihood, likelihood_pairs
from .state import State


def mlp_params(mlp: nn.Module, what: str) -> list[Tensor]:
    """[W1, b1, W2, b2] of a Linear-ELU-Linear head (A6); anything else cannot be fused and raises."""
    ok = (isinstance(mlp, nn.Sequential) and len(mlp) == 3 and isinstance(mlp[0], nn.Linear)  # noqa: PLR2004
          and isinstance(mlp[1], nn.ELU) and isinstance(mlp[2], nn.Linear))
    if not ok:
        msg = f"the fused rollout needs `{what}` to be Linear -> ELU -> Linear (torchrl MLP depth=1, activation ELU), got {mlp}"
        raise RuntimeError(msg)
    return [mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias]


def flat_stoch(stoch: Tensor) -> Tensor:
    """networks.py:162-167: tolerate an un-flattened [B, C, K] stoch."""
    return stoch.flatten(start_dim=1) if stoch.dim() >= 3 else stoch  # noqa: PLR2004


class MoPoE_MRSSM(BaseRSSM):  # noqa: N801
    """Multimodal RSSM with MoPoE posteriors: PoE of {audio, vision}, then MoE of {audio, vision, PoE}."""

    def __init__(  # noqa: PLR0913
        self,
        *,
        audio_representation: Representation,
        vision_representation: Representation,
        transition: Transition,
        audio_encoder: nn.Module,
        vision_encoder: nn.Module,
        audio_decoder: nn.Module,
        vision_decoder: nn.Module,
        init_proj: nn.Module,
        kl_coeff: float,
        use_kl_balancing: bool,
    ) -> None:
        super().__init__(
            representation=audio_representation,  # alias kept: `representation.*` == `audio_representation.*` (:49,:55)
            transition=transition, init_proj=init_proj, kl_coeff=kl_coeff, use_kl_balancing=use_kl_balancing,
        )
        self.audio_representation = audio_representation
        self.vision_representation = vision_representation
        self.audio_encoder = audio_encoder
        self.vision_encoder = vision_encoder
        self.audio_decoder = audio_decoder
        self.vision_decoder = vision_decoder

    # ---- fused-kernel plumbing ------------------------------------------------------------------------------------
    def rollout_weights(self) -> list[Tensor]:
        """Parameters in C-ABI order (`params.MR_STATE_KEYS`)."""
        tr = self.transition
        cell = tr.rnn_cell
        return [
            *mlp_params(tr.action_state_projector, "transition.action_state_projector"),
            cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh,
            *mlp_params(tr.rnn_to_prior_projector, "transition.rnn_to_prior_projector"),
            *mlp_params(self.audio_representation.rnn_to_post_projector, "audio_representation.rnn_to_post_projector"),
            *mlp_params(self.vision_representation.rnn_to_post_projector, "vision_representation.rnn_to_post_projector"),
        ]

    def _class_size(self) -> int:
        return int(self.audio_representation.distribution_factory.class_size)

    # ---- reference API --------------------------------------------------------------------------------------------------
    def encode_observation(self, observation):  # noqa: ANN001, ANN201
        """(:165-182) mean of the two modality embeddings; a Tensor passes through."""
        if isinstance(observation, tuple):
            audio_obs, vision_obs = observation
            return (self.audio_encoder(audio_obs) + self.vision_encoder(vision_obs)) / 2.0
        return observation

    def rollout_representation(self, *, actions: Tensor, observations, prev_state: State) -> tuple[State, State]:  # noqa: ANN001
        """(:184-260) -> (mixed posterior, prior), each stacked over T.  One fused kernel for the whole T loop."""
        if not isinstance(observations, tuple):
            msg = "MoPoE-MRSSM requires tuple of (audio_obs, vision_obs)"
            raise TypeError(msg)
        audio_obs, vision_obs = observations
        audio_embed, vision_embed = self.audio_encoder(audio_obs), self.vision_encoder(vision_obs)
        B, T = audio_embed.shape[:2]
        K = self._class_size()
        C = int(self.audio_representation.distribution_factory.category_size)
        dev = audio_embed.device
        # the reference draws inside State.__init__ from the global RNG (state.py:17); the kernel takes the uniforms
        u_post, u_prior = torch.rand(B, T, C, device=dev), torch.rand(B, T, C, device=dev)
        out = rollout_ops.mrssm_rollout(
            self.rollout_weights(), actions=actions, embed_a=audio_embed, embed_v=vision_embed, h0=prev_state.deter,
            z0=flat_stoch(prev_state.stoch), u_post=u_post, u_prior=u_prior, class_size=K, precision=self._precision(),
            use_kl_balancing=bool(self.use_kl_balancing),
        )
        feature = out["feature"]
        D = prev_state.deter.shape[-1]
        link = FusedKL(kl=out["kl"], use_balancing=bool(self.use_kl_balancing), token=object())
        deter = feature[..., :D]
        posterior = State(deter=deter, stoch=feature[..., D:], feature=feature,
                          distribution=Distribution(out["post_probs"], _fused=link, _role="post"))
        prior = State(deter=deter, stoch=out["prior_stoch"], distribution=Distribution(out["prior_probs"], _fused=link, _role="prior"))
        return posterior, prior

    def rollout_transition(self, *, actions: Tensor, prev_state: State) -> State:
        """(core.py:170-185) imagination; forward-only fused kernel (the reference calls it under no_grad)."""
        if torch.is_grad_enabled() and (actions.requires_grad or prev_state.deter.requires_grad):
            msg = "the fused rollout_transition is forward-only; call it under torch.no_grad() (as the reference's callbacks do)"
            raise RuntimeError(msg)
        B, T = actions.shape[:2]
        fac = self.transition.distribution_factory
        u = torch.rand(B, T, int(fac.category_size), device=actions.device)
        out = rollout_ops.mrssm_imagine(
            [w.detach() for w in self.rollout_weights()], actions=actions, h0=prev_state.deter, z0=flat_stoch(prev_state.stoch), u=u,
            class_size=int(fac.class_size), precision=self._precision(),
        )
        feature = out["feature"]
        D = prev_state.deter.shape[-1]
        return State(deter=feature[..., :D], stoch=feature[..., D:], feature=feature, distribution=Distribution(out["probs"]))

    def decode_state(self, state: State) -> dict[str, Tensor]:
        """(:262-277)"""
        return {"recon/audio": self.audio_decoder(state.feature), "recon/vision": self.vision_decoder(state.feature)}

    @staticmethod
    def compute_reconstruction_loss(reconstructions: dict[str, Tensor], targets: dict[str, Tensor]) -> dict[str, Tensor]:
        """(:279-308)  On CUDA both modalities' likelihoods are ONE launch of the fused streaming kernel (objective.py)."""
        pa, pv, ta, tv = reconstructions["recon/audio"], reconstructions["recon/vision"], targets["recon/audio"], targets["recon/vision"]
        if pa.is_cuda and pv.is_cuda and pa.dtype == pv.dtype:
            audio, vision = likelihood_pairs([pa, pv], [ta, tv], event_ndims=3).unbind(0)
        else:
            audio = likelihood(prediction=pa, target=ta, event_ndims=3)
            vision = likelihood(prediction=pv, target=tv, event_ndims=3)
        return {"recon": audio + vision, "recon/audio": audio, "recon/vision": vision}

    @staticmethod
    def get_observations_from_batch(batch: tuple[Tensor, ...]) -> tuple[Tensor, Tensor]:
        """(:310-323) batch = (action_in, audio_in, vision_in, action_tgt, audio_tgt, vision_tgt)"""
        return batch[1], batch[2]

    @staticmethod
    def get_initial_observation(observations: tuple[Tensor, Tensor]) -> tuple[Tensor, Tensor]:
        """(:325-337)"""
        audio_obs, vision_obs = observations
        return audio_obs[:, 0], vision_obs[:, 0]

    @staticmethod
    def get_targets_from_batch(batch: tuple[Tensor, ...]) -> dict[str, Tensor]:
        """(:339-355)"""
        return {"recon/audio": batch[4], "recon/vision": batch[5]}
