"""`MTRNN`, `MoPoE_MMTRSSM` -- mirror of the reference's `models/mmtrssm/mopoe_mmtrssm/core.py` (:12-610).

Same constructor (incl. the dummy `transition` and the never-used `l_posterior`, which must exist for
`load_state_dict`), module names, `feature_dim`, `initial_state`, rollouts and the hierarchical-KL `shared_step`.
The T loops (:405-490, :512-542) are replaced by the fused CUDA rollout; MTRNN.hidden is passed in and out of the
kernel functionally (the module attribute is still updated afterwards for callers that peek at it)."""

from __future__ import annotations

import torch
from torch import Tensor, nn

from . import _lib, rollout_ops
from .distribution import Distribution, FusedKL, MultiOneHotFactory, kl_divergence
from .mopoe_mrssm import MoPoE_MRSSM, flat_stoch, mlp_params
from .mtstate import MTState
from .networks import Representation, Transition


class MTRNN(nn.Module):
    """Multi-timescale leaky-integrator cell (reference: :12-74): u <- (1-1/tau) u + (W_d d + W_x x)/tau, d = tanh u."""

    def __init__(self, input_dim: int, hidden_dim: int, bias: bool = True, tau: float = 2.0) -> None:
        super().__init__()
        self.hidden_dim = hidden_dim
        self.input_dim = input_dim
        self.tau = tau
        assert tau > 1.0, "tau must be greater than 1.0"
        self._d2h = nn.Linear(hidden_dim, hidden_dim, bias=bias)
        self._input2h = nn.Linear(input_dim, hidden_dim, bias=bias)
        self.hidden: Tensor | None = None

    def forward(self, inputs: Tensor, prev_d: Tensor) -> Tensor:
        if self.hidden is None:
            self.hidden = torch.zeros(inputs.shape[0], self.hidden_dim, device=inputs.device, dtype=inputs.dtype)
        self.hidden = (1 - 1 / self.tau) * self.hidden + (self._d2h(prev_d) + self._input2h(inputs)) / self.tau
        return torch.tanh(self.hidden)


class MoPoE_MMTRSSM(MoPoE_MRSSM):  # noqa: N801
    """MoPoE-MRSSM with a two-level (fast lower / slow higher) multi-timescale latent hierarchy."""

    # bf16 path of this model: one-kernel backward (BPTT + weight gradients on tcgen05 / TMEM), never slower than the
    # two-kernel backward at any batch size measured; "bf16_two_kernel" selects the latter
    _BF16_POLICY = 2  # _lib.PRECISION_BF16_FUSED
    #: SURVEY.md §8 f2 (off by default: results are the reference's up to the rounding of one reassociated sum).  True: under the
    #: bf16 fused policy the embedding half of the modality heads' first layer runs as one fp32 GEMM per modality before the rollout
    #: (`rollout_ops.obs_projection`) and the kernels take the pre-multiplied partials (`obs_projected=True`): half the embedding
    #: bytes at the kernel boundary, no embedding work inside the recurrence.  Parameters and state_dict are unchanged.
    hoist_obs_projection: bool = False

    def __init__(  # noqa: PLR0913
        self,
        *,
        audio_representation: Representation,
        vision_representation: Representation,
        audio_encoder: nn.Module,
        vision_encoder: nn.Module,
        audio_decoder: nn.Module,
        vision_decoder: nn.Module,
        init_proj: nn.Module,
        kl_coeff: float,
        use_kl_balancing: bool,
        action_size: int,
        hd_dim: int,
        hs_dim: int,
        ld_dim: int,
        ls_dim: int,
        l_tau: float,
        h_tau: float,
        l_prior: nn.Module,
        l_posterior: nn.Module,
        h_prior: nn.Module,
        h_posterior: nn.Module,
        l_dist: MultiOneHotFactory,
        h_dist: MultiOneHotFactory,
        w_kl_h: float = 1.0,
    ) -> None:
        dummy_transition = Transition(  # never evaluated; registered so the state_dict matches (:145-151)
            deterministic_size=ld_dim, hidden_size=ld_dim, action_size=1, distribution_config=[1, 1], activation_name="ELU"
        )
        super().__init__(
            audio_representation=audio_representation, vision_representation=vision_representation, transition=dummy_transition,
            audio_encoder=audio_encoder, vision_encoder=vision_encoder, audio_decoder=audio_decoder, vision_decoder=vision_decoder,
            init_proj=init_proj, kl_coeff=kl_coeff, use_kl_balancing=use_kl_balancing,
        )
        self.action_dim = action_size
        self.hd_dim, self.hs_dim, self.ld_dim, self.ls_dim = hd_dim, hs_dim, ld_dim, ls_dim
        self.w_kl_h = w_kl_h
        self.l_rnn = MTRNN(input_dim=action_size + ls_dim + hs_dim, hidden_dim=ld_dim, tau=l_tau)
        self.h_rnn = MTRNN(input_dim=hs_dim, hidden_dim=hd_dim, tau=h_tau)
        self.l_prior, self.l_posterior = l_prior, l_posterior  # l_posterior is never used by the rollout (:405-490)
        self.h_prior, self.h_posterior = h_prior, h_posterior
        self.l_dist, self.h_dist = l_dist, h_dist

    @property
    def feature_dim(self) -> int:
        """(:196-204)"""
        return self.hd_dim + self.hs_dim + self.ld_dim + self.ls_dim

    # ---- fused-kernel plumbing ------------------------------------------------------------------------------------
    def rollout_weights(self) -> list[Tensor]:
        """Parameters in C-ABI order (`params.MT_STATE_KEYS`)."""
        lr, hr = self.l_rnn, self.h_rnn
        return [
            lr._d2h.weight, lr._d2h.bias, lr._input2h.weight, lr._input2h.bias,  # noqa: SLF001
            hr._d2h.weight, hr._d2h.bias, hr._input2h.weight, hr._input2h.bias,  # noqa: SLF001
            *mlp_params(self.l_prior, "l_prior"), *mlp_params(self.h_prior, "h_prior"), *mlp_params(self.h_posterior, "h_posterior"),
            *mlp_params(self.audio_representation.rnn_to_post_projector, "audio_representation.rnn_to_post_projector"),
            *mlp_params(self.vision_representation.rnn_to_post_projector, "vision_representation.rnn_to_post_projector"),
        ]

    def _kernel_cfg(self) -> dict:
        return dict(class_size_l=int(self.l_dist.class_size), class_size_h=int(self.h_dist.class_size), l_tau=float(self.l_rnn.tau),
                    h_tau=float(self.h_rnn.tau), precision=self._precision())

    @staticmethod
    def _state_inputs(prev_state: MTState) -> dict[str, Tensor]:
        return dict(deter_h0=prev_state.deter_h, deter_l0=prev_state.deter_l, hidden_h0=prev_state.hidden_h, hidden_l0=prev_state.hidden_l,
                    stoch_h0=flat_stoch(prev_state.stoch_h), stoch_l0=flat_stoch(prev_state.stoch_l))

    def _split(self, feature: Tensor) -> tuple[Tensor, Tensor, Tensor, Tensor]:
        a, b, c = self.hd_dim, self.hd_dim + self.hs_dim, self.hd_dim + self.hs_dim + self.ld_dim
        return feature[..., :a], feature[..., a:b], feature[..., b:c], feature[..., c:]

    # ---- reference API --------------------------------------------------------------------------------------------------
    def initial_state(self, observation) -> MTState:  # noqa: ANN001
        """(:321-362) init_proj output split into (higher, lower), used raw as deter AND hidden; stoch from the priors."""
        obs_embed = self.encode_observation(observation) if isinstance(observation, tuple) else observation
        h = self.init_proj(obs_embed)
        higher, lower = h[..., : self.hd_dim], h[..., self.hd_dim:]
        self.h_rnn.hidden, self.l_rnn.hidden = higher, lower
        return MTState(
            deter_h=higher, deter_l=lower, distribution_h=self.h_dist(self.h_prior(higher)), distribution_l=self.l_dist(self.l_prior(lower)),
            hidden_h=higher, hidden_l=lower,
        ).to(obs_embed.device)

    def rollout_representation(self, *, actions: Tensor, observations, prev_state: MTState) -> tuple[MTState, MTState]:  # noqa: ANN001
        """(:364-494) -> (mixed posterior, prior) MTStates stacked over T.  One fused kernel for the whole T loop."""
        if not isinstance(observations, tuple):
            msg = "MoPoE-MMTRSSM requires tuple of (audio_obs, vision_obs)"
            raise TypeError(msg)
        audio_obs, vision_obs = observations
        audio_embed, vision_embed = self.audio_encoder(audio_obs), self.vision_encoder(vision_obs)
        B, T = audio_embed.shape[:2]
        dev = audio_embed.device
        CL, CH = int(self.l_dist.category_size), int(self.h_dist.category_size)
        u = {k: torch.rand(B, T, c, device=dev) for k, c in (("u_post_l", CL), ("u_post_h", CH), ("u_prior_l", CL), ("u_prior_h", CH))}
        cfg = self._kernel_cfg()
        projected = bool(self.hoist_obs_projection) and cfg["precision"] == _lib.PRECISION_BF16_FUSED
        if projected:
            # the embedding half of the two modality heads' first layer (:259-260, `cat([d_l, e]) @ W1.T` =
            # `d_l @ W1[:, :ld].T + e @ W1[:, ld:].T`) as ONE GEMM per modality over all (b,t) before the loop; same parameters,
            # same state_dict; autograd of this GEMM yields d embed and the W1[:, ld:] columns of d W1
            w1a = self.audio_representation.rnn_to_post_projector[0].weight
            w1v = self.vision_representation.rnn_to_post_projector[0].weight
            audio_embed = rollout_ops.obs_projection(audio_embed, w1a, self.ld_dim)
            vision_embed = rollout_ops.obs_projection(vision_embed, w1v, self.ld_dim)
        out = rollout_ops.mtrssm_rollout(
            self.rollout_weights(), actions=actions, embed_a=audio_embed, embed_v=vision_embed, **self._state_inputs(prev_state), **u,
            use_kl_balancing=bool(self.use_kl_balancing), obs_projected=projected, **cfg,
        )
        feature = out["feature"]
        deter_h, stoch_h, deter_l, stoch_l = self._split(feature)
        bal = bool(self.use_kl_balancing)
        link_l = FusedKL(kl=out["kl_l"], use_balancing=bal, token=object())
        link_h = FusedKL(kl=out["kl_h"], use_balancing=bal, token=object())
        posterior = MTState(
            deter_h=deter_h, deter_l=deter_l, stoch_h=stoch_h, stoch_l=stoch_l, feature=feature, hidden_h=out["hidden_h"], hidden_l=out["hidden_l"],
            distribution_h=Distribution(out["post_probs_h"], _fused=link_h, _role="post"),
            distribution_l=Distribution(out["post_probs_l"], _fused=link_l, _role="post"),
        )
        prior = MTState(
            deter_h=deter_h, deter_l=deter_l, stoch_h=out["prior_stoch_h"], stoch_l=out["prior_stoch_l"], hidden_h=out["hidden_h"],
            hidden_l=out["hidden_l"],
            distribution_h=Distribution(out["prior_probs_h"], _fused=link_h, _role="prior"),
            distribution_l=Distribution(out["prior_probs_l"], _fused=link_l, _role="prior"),
        )
        self.h_rnn.hidden, self.l_rnn.hidden = out["hidden_h"][:, -1], out["hidden_l"][:, -1]
        return posterior, prior

    def rollout_transition(self, *, actions: Tensor, prev_state: MTState) -> MTState:  # type: ignore[override]
        """(:496-544) imagination with the priors' own samples fed back; forward-only fused kernel."""
        if torch.is_grad_enabled() and (actions.requires_grad or prev_state.deter_l.requires_grad):
            msg = "the fused rollout_transition is forward-only; call it under torch.no_grad() (as the reference's callbacks do)"
            raise RuntimeError(msg)
        B, T = actions.shape[:2]
        dev = actions.device
        out = rollout_ops.mtrssm_imagine(
            [w.detach() for w in self.rollout_weights()], actions=actions, **self._state_inputs(prev_state),
            u_l=torch.rand(B, T, int(self.l_dist.category_size), device=dev), u_h=torch.rand(B, T, int(self.h_dist.category_size), device=dev),
            **self._kernel_cfg(),
        )
        feature = out["feature"]
        deter_h, stoch_h, deter_l, stoch_l = self._split(feature)
        self.h_rnn.hidden, self.l_rnn.hidden = out["hidden_h"][:, -1], out["hidden_l"][:, -1]
        return MTState(
            deter_h=deter_h, deter_l=deter_l, stoch_h=stoch_h, stoch_l=stoch_l, feature=feature, hidden_h=out["hidden_h"], hidden_l=out["hidden_l"],
            distribution_h=Distribution(out["probs_h"]), distribution_l=Distribution(out["probs_l"]),
        )

    def shared_step(self, batch: tuple[Tensor, ...]) -> dict[str, Tensor]:
        """(:563-606) loss = recon + kl_coeff * KL_l + kl_coeff * w_kl_h * KL_h"""
        observations = self.get_observations_from_batch(batch)
        posterior, prior = self.rollout_representation(
            actions=batch[0], observations=observations, prev_state=self.initial_state(self.get_initial_observation(observations))
        )
        loss_dict = self.compute_reconstruction_loss(self.decode_state(posterior), self.get_targets_from_batch(batch))
        kl_l = kl_divergence(
            q=posterior.distribution_l.independent(1), p=prior.distribution_l.independent(1), use_balancing=self.use_kl_balancing
        ).mul(self.kl_coeff)
        kl_h = kl_divergence(
            q=posterior.distribution_h.independent(1), p=prior.distribution_h.independent(1), use_balancing=self.use_kl_balancing
        ).mul(self.kl_coeff * self.w_kl_h)
        loss_dict["kl"] = kl_l
        loss_dict["kl_h"] = kl_h
        loss_dict["loss"] = loss_dict["recon"] + kl_l + kl_h
        return loss_dict
