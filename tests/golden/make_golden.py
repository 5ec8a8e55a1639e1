"""Generate golden vectors by running the REFERENCE'S OWN hot-path code (test infrastructure).

Run here (the container that has `/root/reference`); the GPU box never runs this:

    python tests/golden/make_golden.py

It imports `/root/reference/src/multimodal_rssm/models/{core,networks,state,objective}.py`,
`.../mrssm/mopoe_mrssm/core.py`, `.../mmtrssm/{state.py,mopoe_mmtrssm/core.py}` UNMODIFIED on top
of the third-party stand-ins in `ref_shims.py` (the only restated pieces: A1..A7 of SURVEY.md
§8(c)), builds both models at the `default.yaml` dims, and records for fixed seeds

* every rollout input (actions, encoder embeddings, initial state, the uniforms each
  `rsample()` consumed, in call order),
* every rollout output (posterior / prior deter, probs, stoch, hidden, feature),
* the `shared_step` loss dict and the gradients it induces on every rollout parameter, on the
  embeddings, on the initial state, plus d(loss)/d(posterior.feature) so a rollout-only
  implementation can replay the exact upstream gradient,
* `rollout_transition` (imagination) outputs started from `posterior[:, -1]`.

Encoders/decoders: `cnn.Encoder/Decoder` are absent; shape-compatible stand-ins (A7) are used and
their outputs are recorded, so the fixtures do not depend on them.

Outputs: `tests/golden/mrssm_default.pt`, `tests/golden/mtrssm_default.pt` (B = 5, T = 7; a few hundred KB),
`mrssm_cfg1.pt`, `mtrssm_cfg2.pt` (the batch and sequence length of the two `default.yaml`s, BASELINE.json configs[0] and [1]:
B = 8, T = 30 -- `mopoe_mrssm/configs/default.yaml:162,180-182`, `mopoe_mmtrssm/configs/default.yaml:209,227-229`; the
observation batch is regenerated from its seed by `tests/helpers.golden_batch` and checked against a stored checksum) and
`rssm_unimodal.pt` (`BaseRSSM.rollout_representation`, `models/core.py:137-168`, reached through a minimal
concrete subclass of the reference's own `BaseRSSM` -- neither shipped model calls it, both override it).
"""

from __future__ import annotations

import sys
from pathlib import Path

import torch
from torch import Tensor, nn

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import ref_shims  # noqa: E402
from ref_shims import MLP, NOISE, MultiOneHotFactory  # noqa: E402

E = 64  # obs_embed_size (default.yaml:10)


class RecEncoder(nn.Module):
    """A7 stand-in for cnn.Encoder: [*,1,32,32] -> [*,64]; records its output (and its grad)."""

    def __init__(self) -> None:
        super().__init__()
        self.lin = nn.Linear(32 * 32, E)
        self.outputs: list[Tensor] = []

    def forward(self, x: Tensor) -> Tensor:
        out = self.lin(x.flatten(start_dim=-3))
        if out.requires_grad:
            out.retain_grad()
        self.outputs.append(out)
        return out


class Decoder(nn.Module):
    """A7 stand-in for cnn.Decoder: [*,F] -> [*,1,32,32], Tanh output."""

    def __init__(self, feature: int) -> None:
        super().__init__()
        self.lin = nn.Linear(feature, 32 * 32)

    def forward(self, x: Tensor) -> Tensor:
        return torch.tanh(self.lin(x)).reshape(*x.shape[:-1], 1, 32, 32)


def synth_batch(B: int, T: int, seed: int) -> tuple[Tensor, ...]:
    """SURVEY.md §8(d) synthetic inputs: one-hot speaker id + N(0,0.1^2); obs ~ U(-1,1)."""
    g = torch.Generator().manual_seed(seed)
    speaker = torch.randint(0, 6, (B,), generator=g)
    act = torch.nn.functional.one_hot(speaker, 6).float()[:, None, :].expand(B, T, 6)
    act_in = act + 0.1 * torch.randn(B, T, 6, generator=g)
    audio = torch.rand(B, T, 1, 32, 32, generator=g) * 2 - 1
    vision = torch.rand(B, T, 1, 32, 32, generator=g) * 2 - 1
    return (act_in, audio, vision, act.clone(), audio.clone(), vision.clone())


def rollout_params(model: nn.Module, prefixes: tuple[str, ...]) -> dict[str, Tensor]:
    # remove_duplicate=False: `representation.*` aliases `audio_representation.*` (mopoe_mrssm/core.py:49,55)
    return {k: v for k, v in model.named_parameters(remove_duplicate=False) if k.startswith(prefixes)}


def grads_of(params: dict[str, Tensor]) -> dict[str, Tensor]:
    return {k: (torch.zeros_like(v) if v.grad is None else v.grad.clone()) for k, v in params.items()}


def build_mrssm(ref) -> nn.Module:  # noqa: ANN001
    """mopoe_mrssm/configs/default.yaml:5-101 with A7 encoder/decoder stand-ins."""
    torch.manual_seed(42)
    rep = dict(deterministic_size=32, hidden_size=32, obs_embed_size=E, distribution_config=[4, 4], activation_name="ELU")
    return ref.mopoe_mrssm.MoPoE_MRSSM(
        audio_representation=ref.networks.Representation(**rep),
        vision_representation=ref.networks.Representation(**rep),
        transition=ref.networks.Transition(
            deterministic_size=32, hidden_size=32, action_size=6, distribution_config=[4, 4], activation_name="ELU"
        ),
        audio_encoder=RecEncoder(),
        vision_encoder=RecEncoder(),
        audio_decoder=Decoder(48),
        vision_decoder=Decoder(48),
        init_proj=MLP(in_features=E, out_features=32, num_cells=200, depth=1),
        kl_coeff=1,
        use_kl_balancing=True,
    )


def build_mtrssm(ref) -> nn.Module:  # noqa: ANN001
    """mopoe_mmtrssm/configs/default.yaml:5-148 with A7 encoder/decoder stand-ins."""
    torch.manual_seed(42)
    rep = dict(deterministic_size=32, hidden_size=32, obs_embed_size=E, distribution_config=[4, 4], activation_name="ELU")

    def head(i: int) -> nn.Module:
        return MLP(in_features=i, out_features=16, num_cells=32, depth=1, activation_class=nn.ELU)

    return ref.mopoe_mmtrssm.MoPoE_MMTRSSM(
        audio_representation=ref.networks.Representation(**rep),
        vision_representation=ref.networks.Representation(**rep),
        audio_encoder=RecEncoder(),
        vision_encoder=RecEncoder(),
        audio_decoder=Decoder(96),
        vision_decoder=Decoder(96),
        init_proj=MLP(in_features=E, out_features=64, num_cells=200, depth=1),
        kl_coeff=1,
        use_kl_balancing=True,
        action_size=6,
        hd_dim=32,
        hs_dim=16,
        ld_dim=32,
        ls_dim=16,
        l_tau=2.0,
        h_tau=4.0,
        l_prior=head(32),
        l_posterior=head(96),
        h_prior=head(32),
        h_posterior=head(64),
        l_dist=MultiOneHotFactory(class_size=4, category_size=4),
        h_dist=MultiOneHotFactory(class_size=2, category_size=8),
        w_kl_h=1.0,
    )


def pack_batch(batch: tuple[Tensor, ...], seed: int, compact: bool) -> dict:
    """Small fixtures keep the batch; the default.yaml-sized ones keep its seed and a checksum (synth_batch is deterministic)."""
    if not compact:
        return {"batch": tuple(t.clone() for t in batch)}
    return {"batch_seed": seed, "batch_checksum": [float(t.double().sum()) for t in batch]}


def golden_mrssm(ref, B: int, T: int, Ti: int, compact: bool = False) -> dict:  # noqa: ANN001
    model = build_mrssm(ref)
    batch = synth_batch(B, T, seed=1234)
    obs = model.get_observations_from_batch(batch)

    # (1) the reference's shared_step, noise recorded
    NOISE.reset(4321)
    loss_ref = model.shared_step(batch)
    noise_log = list(NOISE.log)

    # (2) the same computation, piece by piece with the recorded noise, to expose intermediates
    model.zero_grad()
    for enc in (model.audio_encoder, model.vision_encoder):
        enc.outputs.clear()
    NOISE.reset(0)
    NOISE.forced.extend(noise_log)
    init = model.initial_state(model.get_initial_observation(obs))
    init.deter.retain_grad()
    init.stoch.retain_grad()
    post, prior = model.rollout_representation(actions=batch[0], observations=obs, prev_state=init)
    post.feature.retain_grad()
    recon = model.decode_state(post)
    loss = model.compute_reconstruction_loss(recon, model.get_targets_from_batch(batch))
    from distribution_extension import kl_divergence

    kl = kl_divergence(
        q=post.distribution.independent(1), p=prior.distribution.independent(1), use_balancing=model.use_kl_balancing
    ).mul(model.kl_coeff)
    total = loss["recon"] + kl
    assert torch.equal(total, loss_ref["loss"]), (total, loss_ref["loss"])
    assert torch.equal(kl, loss_ref["kl"])
    total.backward()

    # noise order (models/state.py:17 via networks.py:173, mopoe_mrssm/core.py:83,163):
    # initial z0, then per step: prior, audio (discarded), vision (discarded), mixed posterior
    assert len(noise_log) == 1 + 4 * T
    u_prior = torch.stack([noise_log[1 + 4 * t + 0] for t in range(T)], 1)
    u_post = torch.stack([noise_log[1 + 4 * t + 3] for t in range(T)], 1)

    # encoder call order: initial_state (audio, vision) on [B,...], then rollout (audio, vision) on [B,T,...]
    ea, ev = model.audio_encoder.outputs[1], model.vision_encoder.outputs[1]
    params = rollout_params(model, ("transition.", "audio_representation.", "vision_representation."))

    # (3) imagination from posterior[:, -1] (mrssm/callback.py:184-188 usage)
    with torch.no_grad():
        NOISE.reset(777)
        act_im = synth_batch(B, Ti, seed=99)[0]
        imag = model.rollout_transition(actions=act_im, prev_state=post[:, -1])
        u_imag = torch.stack(list(NOISE.log), 1)

    full_grads = {k: (torch.zeros_like(v) if v.grad is None else v.grad.clone()) for k, v in model.named_parameters()}
    return {
        "full_state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
        "full_grads": full_grads,
        **pack_batch(batch, 1234, compact),
        "dims": dict(B=B, T=T, Ti=Ti, A=6, E=E, D=32, H=32, C=4, K=4, kl_coeff=1.0, use_kl_balancing=True),
        "params": {k: v.detach().clone() for k, v in params.items()},
        "inputs": {
            "actions": batch[0].clone(),
            "embed_a": ea.detach().clone(),
            "embed_v": ev.detach().clone(),
            "h0": init.deter.detach().clone(),
            "z0": init.stoch.detach().clone(),
            "u_z0": noise_log[0],
            "u_prior": u_prior,
            "u_post": u_post,
        },
        "outputs": {
            "deter": post.deter.detach().clone(),
            "post_probs": post.distribution.probs.detach().clone(),
            "post_stoch": post.stoch.detach().clone(),
            "post_feature": post.feature.detach().clone(),
            "prior_probs": prior.distribution.probs.detach().clone(),
            "prior_stoch": prior.stoch.detach().clone(),
            "prior_deter": prior.deter.detach().clone(),
        },
        "loss": {k: v.detach().clone() for k, v in loss_ref.items()},
        "upstream": {"d_post_feature": post.feature.grad.clone()},
        "grads": {
            "params": grads_of(params),
            "embed_a": ea.grad.clone(),
            "embed_v": ev.grad.clone(),
            "h0": init.deter.grad.clone(),
            "z0": init.stoch.grad.clone(),
        },
        "imagine": {
            "actions": act_im,
            "u": u_imag,
            "deter": imag.deter.clone(),
            "probs": imag.distribution.probs.clone(),
            "stoch": imag.stoch.clone(),
        },
    }


def golden_mtrssm(ref, B: int, T: int, Ti: int, compact: bool = False) -> dict:  # noqa: ANN001
    model = build_mtrssm(ref)
    batch = synth_batch(B, T, seed=1234)
    obs = model.get_observations_from_batch(batch)

    NOISE.reset(4321)
    loss_ref = model.shared_step(batch)
    noise_log = list(NOISE.log)

    model.zero_grad()
    for enc in (model.audio_encoder, model.vision_encoder):
        enc.outputs.clear()
    NOISE.reset(0)
    NOISE.forced.extend(noise_log)
    init = model.initial_state(model.get_initial_observation(obs))
    # initial_state: deter_* and hidden_* are the SAME tensors (mopoe_mmtrssm/core.py:354-361)
    for t in (init.deter_h, init.deter_l, init.stoch_h, init.stoch_l):
        t.retain_grad()
    post, prior = model.rollout_representation(actions=batch[0], observations=obs, prev_state=init)
    post.feature.retain_grad()
    recon = model.decode_state(post)
    # the class attribute at mopoe_mmtrssm/core.py:609 rebinds a staticmethod as an instance method,
    # so (like shared_step, :584) call it through MoPoE_MRSSM
    loss = ref.mopoe_mrssm.MoPoE_MRSSM.compute_reconstruction_loss(recon, model.get_targets_from_batch(batch))
    from distribution_extension import kl_divergence

    kl_l = kl_divergence(
        q=post.distribution_l.independent(1), p=prior.distribution_l.independent(1), use_balancing=True
    ).mul(model.kl_coeff)
    kl_h = kl_divergence(
        q=post.distribution_h.independent(1), p=prior.distribution_h.independent(1), use_balancing=True
    ).mul(model.kl_coeff * model.w_kl_h)
    total = loss["recon"] + kl_l + kl_h
    assert torch.equal(total, loss_ref["loss"]), (total, loss_ref["loss"])
    total.backward()

    # noise order: initial (h, l) (mmtrssm/state.py:48-49), then per step: l_post (core.py:456),
    # h_post (:464), prior MTState ctor h then l (:467-474 -> state.py:48-49)
    assert len(noise_log) == 2 + 4 * T
    pick = lambda j: torch.stack([noise_log[2 + 4 * t + j] for t in range(T)], 1)  # noqa: E731
    ea, ev = model.audio_encoder.outputs[1], model.vision_encoder.outputs[1]
    params = rollout_params(
        model,
        ("l_rnn.", "h_rnn.", "l_prior.", "h_prior.", "h_posterior.", "audio_representation.", "vision_representation."),
    )
    # parameters that never receive gradients: dummy transition, l_posterior
    dead = {k: v.grad is None for k, v in model.named_parameters() if k.startswith(("transition.", "l_posterior."))}
    assert all(dead.values()), dead

    with torch.no_grad():
        NOISE.reset(777)
        act_im = synth_batch(B, Ti, seed=99)[0]
        imag = model.rollout_transition(actions=act_im, prev_state=post[:, -1])
        log = list(NOISE.log)  # per step: prior h then l
        u_imag_h = torch.stack(log[0::2], 1)
        u_imag_l = torch.stack(log[1::2], 1)

    full_grads = {k: (torch.zeros_like(v) if v.grad is None else v.grad.clone()) for k, v in model.named_parameters()}
    return {
        "full_state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
        "full_grads": full_grads,
        **pack_batch(batch, 1234, compact),
        "dims": dict(
            B=B, T=T, Ti=Ti, A=6, E=E, HD=32, LD=32, HR=32, HH=32, CL=4, KL=4, CH=8, KH=2,
            l_tau=2.0, h_tau=4.0, kl_coeff=1.0, w_kl_h=1.0, use_kl_balancing=True,
        ),
        "params": {k: v.detach().clone() for k, v in params.items()},
        "inputs": {
            "actions": batch[0].clone(),
            "embed_a": ea.detach().clone(),
            "embed_v": ev.detach().clone(),
            "deter_h0": init.deter_h.detach().clone(),
            "deter_l0": init.deter_l.detach().clone(),
            "hidden_h0": init.hidden_h.detach().clone(),
            "hidden_l0": init.hidden_l.detach().clone(),
            "stoch_h0": init.stoch_h.detach().clone(),
            "stoch_l0": init.stoch_l.detach().clone(),
            "u_h0": noise_log[0],
            "u_l0": noise_log[1],
            "u_post_l": pick(0),
            "u_post_h": pick(1),
            "u_prior_h": pick(2),
            "u_prior_l": pick(3),
        },
        "outputs": {
            "deter_h": post.deter_h.detach().clone(),
            "deter_l": post.deter_l.detach().clone(),
            "hidden_h": post.hidden_h.detach().clone(),
            "hidden_l": post.hidden_l.detach().clone(),
            "post_probs_h": post.distribution_h.probs.detach().clone(),
            "post_probs_l": post.distribution_l.probs.detach().clone(),
            "post_stoch_h": post.stoch_h.detach().clone(),
            "post_stoch_l": post.stoch_l.detach().clone(),
            "post_feature": post.feature.detach().clone(),
            "prior_probs_h": prior.distribution_h.probs.detach().clone(),
            "prior_probs_l": prior.distribution_l.probs.detach().clone(),
            "prior_stoch_h": prior.stoch_h.detach().clone(),
            "prior_stoch_l": prior.stoch_l.detach().clone(),
        },
        "loss": {k: v.detach().clone() for k, v in loss_ref.items()},
        "upstream": {"d_post_feature": post.feature.grad.clone()},
        "grads": {
            "params": grads_of(params),
            "embed_a": ea.grad.clone(),
            "embed_v": ev.grad.clone(),
            # deter_*0 and hidden_*0 alias one tensor in the reference, so this is the SUM of both paths
            "deter_hidden_h0": init.deter_h.grad.clone(),
            "deter_hidden_l0": init.deter_l.grad.clone(),
            "stoch_h0": init.stoch_h.grad.clone(),
            "stoch_l0": init.stoch_l.grad.clone(),
        },
        "imagine": {
            "actions": act_im,
            "u_h": u_imag_h,
            "u_l": u_imag_l,
            "deter_h": imag.deter_h.clone(),
            "deter_l": imag.deter_l.clone(),
            "hidden_h": imag.hidden_h.clone(),
            "hidden_l": imag.hidden_l.clone(),
            "probs_h": imag.distribution_h.probs.clone(),
            "probs_l": imag.distribution_l.probs.clone(),
            "stoch_h": imag.stoch_h.clone(),
            "stoch_l": imag.stoch_l.clone(),
        },
    }


def golden_unimodal(ref, B: int, T: int, Ti: int) -> dict:  # noqa: ANN001
    """the reference's OWN `BaseRSSM.rollout_representation` / `initial_state` / `shared_step` /
    `rollout_transition` (models/core.py:121-221), reached through the smallest concrete subclass: one modality (the audio
    tensors of the batch), the hooks filled in the way MoPoE_MRSSM fills them (mopoe_mrssm/core.py:165-182,262-355)."""
    torch.manual_seed(42)

    class Unimodal(ref.core.BaseRSSM):
        def __init__(self, *, encoder: nn.Module, decoder: nn.Module, **kw) -> None:  # noqa: ANN003
            super().__init__(**kw)
            self.encoder, self.decoder = encoder, decoder

        def encode_observation(self, observation: Tensor) -> Tensor:
            return self.encoder(observation)

        def decode_state(self, state) -> dict[str, Tensor]:  # noqa: ANN001
            return {"recon": self.decoder(state.feature)}

        def compute_reconstruction_loss(self, reconstructions, targets) -> dict[str, Tensor]:  # noqa: ANN001
            return {"recon": ref.objective.likelihood(prediction=reconstructions["recon"], target=targets["recon"], event_ndims=3)}

        def get_observations_from_batch(self, batch):  # noqa: ANN001, ANN201
            return batch[1]

        def get_initial_observation(self, observations):  # noqa: ANN001, ANN201
            return observations[:, 0]

        def get_targets_from_batch(self, batch):  # noqa: ANN001, ANN201
            return {"recon": batch[4]}

    dist = [4, 4]
    model = Unimodal(
        representation=ref.networks.Representation(deterministic_size=32, hidden_size=32, obs_embed_size=E, distribution_config=dist, activation_name="ELU"),
        transition=ref.networks.Transition(deterministic_size=32, hidden_size=32, action_size=6, distribution_config=dist, activation_name="ELU"),
        init_proj=MLP(in_features=E, out_features=32, num_cells=200, depth=1), kl_coeff=1, use_kl_balancing=True,
        encoder=RecEncoder(), decoder=Decoder(48),
    )
    batch = synth_batch(B, T, seed=1234)
    NOISE.reset(4321)
    loss_ref = model.shared_step(batch)
    noise_log = list(NOISE.log)

    model.zero_grad()
    model.encoder.outputs.clear()
    NOISE.reset(0)
    NOISE.forced.extend(noise_log)
    obs = model.get_observations_from_batch(batch)
    init = model.initial_state(model.get_initial_observation(obs))
    init.deter.retain_grad()
    init.stoch.retain_grad()
    post, prior = model.rollout_representation(actions=batch[0], observations=obs, prev_state=init)  # core.py:137-168
    post.feature.retain_grad()
    loss = model.compute_reconstruction_loss(model.decode_state(post), model.get_targets_from_batch(batch))
    from distribution_extension import kl_divergence

    kl = kl_divergence(q=post.distribution.independent(1), p=prior.distribution.independent(1), use_balancing=True).mul(model.kl_coeff)
    total = loss["recon"] + kl
    assert torch.equal(total, loss_ref["loss"]), (total, loss_ref["loss"])
    total.backward()
    # noise order: initial z0, then per step the prior State (networks.py:173) and the posterior State (networks.py:84)
    assert len(noise_log) == 1 + 2 * T
    u_prior = torch.stack([noise_log[1 + 2 * t] for t in range(T)], 1)
    u_post = torch.stack([noise_log[2 + 2 * t] for t in range(T)], 1)
    embed = model.encoder.outputs[1]
    params = rollout_params(model, ("transition.", "representation."))
    with torch.no_grad():
        NOISE.reset(777)
        act_im = synth_batch(B, Ti, seed=99)[0]
        imag = model.rollout_transition(actions=act_im, prev_state=post[:, -1])  # core.py:170-185
        u_imag = torch.stack(list(NOISE.log), 1)
    return {
        "full_state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
        "full_grads": {k: (torch.zeros_like(v) if v.grad is None else v.grad.clone()) for k, v in model.named_parameters()},
        "batch": tuple(t.clone() for t in batch),
        "dims": dict(B=B, T=T, Ti=Ti, A=6, E=E, D=32, H=32, C=4, K=4, kl_coeff=1.0, use_kl_balancing=True),
        "params": {k: v.detach().clone() for k, v in params.items()},
        "inputs": {"actions": batch[0].clone(), "embed": embed.detach().clone(), "h0": init.deter.detach().clone(),
                   "z0": init.stoch.detach().clone(), "u_z0": noise_log[0], "u_prior": u_prior, "u_post": u_post},
        "outputs": {"deter": post.deter.detach().clone(), "post_probs": post.distribution.probs.detach().clone(),
                    "post_stoch": post.stoch.detach().clone(), "post_feature": post.feature.detach().clone(),
                    "prior_probs": prior.distribution.probs.detach().clone(), "prior_stoch": prior.stoch.detach().clone(),
                    "prior_deter": prior.deter.detach().clone()},
        "loss": {k: v.detach().clone() for k, v in loss_ref.items()},
        "upstream": {"d_post_feature": post.feature.grad.clone()},
        "grads": {"params": grads_of(params), "embed": embed.grad.clone(), "h0": init.deter.grad.clone(), "z0": init.stoch.grad.clone()},
        "imagine": {"actions": act_im, "u": u_imag, "deter": imag.deter.clone(), "probs": imag.distribution.probs.clone(),
                    "stoch": imag.stoch.clone()},
    }


def main() -> None:
    torch.set_num_threads(1)
    torch.use_deterministic_algorithms(True)
    ref = ref_shims.import_reference()
    g1 = golden_mrssm(ref, B=5, T=7, Ti=4)
    torch.save(g1, HERE / "mrssm_default.pt")
    g2 = golden_mtrssm(ref, B=5, T=7, Ti=4)
    torch.save(g2, HERE / "mtrssm_default.pt")
    # the two default.yaml configurations at their own batch size and sequence length (BASELINE.json configs[0], configs[1])
    g3 = golden_mrssm(ref, B=8, T=30, Ti=10, compact=True)
    torch.save(g3, HERE / "mrssm_cfg1.pt")
    g4 = golden_mtrssm(ref, B=8, T=30, Ti=10, compact=True)
    torch.save(g4, HERE / "mtrssm_cfg2.pt")
    g5 = golden_unimodal(ref, B=6, T=9, Ti=4)
    torch.save(g5, HERE / "rssm_unimodal.pt")
    for name, g in (("mrssm", g1), ("mtrssm", g2), ("mrssm cfg1", g3), ("mtrssm cfg2", g4), ("unimodal", g5)):
        print(name, {k: float(v) for k, v in g["loss"].items()})


if __name__ == "__main__":
    main()
