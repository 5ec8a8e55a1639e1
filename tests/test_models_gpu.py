"""GPU tests of the drop-in model classes: the product's `shared_step` / rollouts (fused CUDA kernels behind the
reference's Python interface) against golden vectors produced by the reference's own `shared_step`."""

import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _tol(ref):
    return dict(rtol=1e-4, atol=2e-5 * max(float(ref.abs().max()), 1e-3))


@pytest.mark.parametrize("fixture", H.MRSSM_GOLDEN)
def test_mrssm_shared_step_matches_reference_losses_and_all_gradients(golden_dir, fixture):
    """Drop-in check: same weights (state_dict), same batch, same noise -> same loss dict and the same gradient on
    EVERY parameter (encoders, decoders, init_proj and the rollout), including the initial_state paths."""
    g = torch.load(golden_dir / fixture)
    model = H.build_mrssm_model()
    model.load_state_dict(g["full_state_dict"], strict=True)
    model.cuda()
    inp = g["inputs"]
    batch = tuple(t.cuda() for t in H.golden_batch(g))
    # product noise order: initial_state draw, then the rollout's u_post, u_prior (mopoe_mrssm.py)
    with H.RandQueue([inp["u_z0"], inp["u_post"], inp["u_prior"]]) as q:
        loss = model.shared_step(batch)
        assert not q.values
    rep = H.Report("MoPoE_MRSSM.shared_step vs the reference's shared_step")
    for k, v in g["loss"].items():
        rep.check(k, loss[k], v, rtol=1e-5, atol=1e-6)
    loss["loss"].backward()
    for k, p in model.named_parameters():
        rep.check("d " + k, p.grad if p.grad is not None else torch.zeros_like(p), g["full_grads"][k], **_tol(g["full_grads"][k]))
    rep.finish()


@pytest.mark.parametrize("fixture", H.MTRSSM_GOLDEN)
def test_mtrssm_shared_step_matches_reference_losses_and_all_gradients(golden_dir, fixture):
    g = torch.load(golden_dir / fixture)
    model = H.build_mtrssm_model()
    model.load_state_dict(g["full_state_dict"], strict=True)
    model.cuda()
    inp = g["inputs"]
    batch = tuple(t.cuda() for t in H.golden_batch(g))
    # initial_state draws h then l (mmtrssm/state.py:48-49); the rollout draws post_l, post_h, prior_l, prior_h
    noise = [inp["u_h0"], inp["u_l0"], inp["u_post_l"], inp["u_post_h"], inp["u_prior_l"], inp["u_prior_h"]]
    with H.RandQueue(noise) as q:
        loss = model.shared_step(batch)
        assert not q.values
    rep = H.Report("MoPoE_MMTRSSM.shared_step vs the reference's shared_step")
    for k, v in g["loss"].items():
        rep.check(k, loss[k], v, rtol=1e-5, atol=1e-6)
    loss["loss"].backward()
    dead = 0
    for k, p in model.named_parameters():
        if p.grad is None:  # the reference leaves these without gradient too
            assert k.startswith(("transition.", "l_posterior.")), k
            assert not g["full_grads"][k].any()
            dead += 1
            continue
        rep.check("d " + k, p.grad, g["full_grads"][k], **_tol(g["full_grads"][k]))
    assert dead > 0
    rep.finish()


def test_rollout_transition_matches_reference_imagination(golden_dir):
    """callbacks' usage (mrssm/callback.py:184-188): imagination from posterior[:, -1]."""
    for name, build in (("mrssm_default.pt", H.build_mrssm_model), ("mtrssm_default.pt", H.build_mtrssm_model),
                        ("mrssm_cfg1.pt", H.build_mrssm_model), ("mtrssm_cfg2.pt", H.build_mtrssm_model)):
        g = torch.load(golden_dir / name)
        model = build()
        model.load_state_dict(g["full_state_dict"], strict=True)
        model.cuda()
        im, out = g["imagine"], g["outputs"]
        rep = H.Report(f"{type(model).__name__}.rollout_transition vs reference")
        with torch.no_grad():
            if "deter" in out:
                from multimodal_mtrssm_b200.distribution import Distribution
                from multimodal_mtrssm_b200.state import State

                prev = State(deter=out["deter"][:, -1].cuda(), stoch=out["post_stoch"][:, -1].cuda(),
                             distribution=Distribution(out["post_probs"][:, -1].cuda()))
                with H.RandQueue([im["u"]]):
                    got = model.rollout_transition(actions=im["actions"].cuda(), prev_state=prev)
                rep.check("deter", got.deter, im["deter"], rtol=1e-5, atol=2e-6)
                rep.check("stoch", got.stoch, im["stoch"], rtol=1e-5, atol=2e-6)
                rep.check("probs", got.distribution.probs, im["probs"], rtol=1e-5, atol=2e-6)
                assert got.feature.shape == (*im["deter"].shape[:2], 48)
            else:
                from multimodal_mtrssm_b200.distribution import Distribution
                from multimodal_mtrssm_b200.mtstate import MTState

                last = lambda k: out[k][:, -1].cuda()  # noqa: E731
                prev = MTState(deter_h=last("deter_h"), deter_l=last("deter_l"), hidden_h=last("hidden_h"), hidden_l=last("hidden_l"),
                               stoch_h=last("post_stoch_h"), stoch_l=last("post_stoch_l"),
                               distribution_h=Distribution(last("post_probs_h")), distribution_l=Distribution(last("post_probs_l")))
                with H.RandQueue([im["u_l"], im["u_h"]]):
                    got = model.rollout_transition(actions=im["actions"].cuda(), prev_state=prev)
                for k in ("deter_h", "deter_l", "hidden_h", "hidden_l", "stoch_h", "stoch_l"):
                    rep.check(k, getattr(got, k), im[k], rtol=1e-5, atol=2e-6)
                rep.check("probs_h", got.distribution_h.probs, im["probs_h"], rtol=1e-5, atol=2e-6)
                rep.check("probs_l", got.distribution_l.probs, im["probs_l"], rtol=1e-5, atol=2e-6)
        rep.finish()


def test_fused_states_behave_like_reference_states():
    """Consumers slice / concatenate the stacked States (mrssm/callback.py:184-188,222-225)."""
    from multimodal_mtrssm_b200.distribution import kl_divergence
    from multimodal_mtrssm_b200.state import cat_states

    torch.manual_seed(0)
    model = H.build_mrssm_model().cuda()
    B, T = 6, 9
    batch = tuple(t.cuda() for t in (torch.randn(B, T, 6), torch.rand(B, T, 1, 32, 32), torch.rand(B, T, 1, 32, 32)))
    obs = (batch[1], batch[2])
    post, prior = model.rollout_representation(actions=batch[0], observations=obs, prev_state=model.initial_state((obs[0][:, 0], obs[1][:, 0])))
    assert post.feature.shape == (B, T, 48) and post.deter.data_ptr() == post.feature.data_ptr()  # views of one tensor
    assert torch.equal(post.deter, prior.deter)  # posterior reuses the prior's deter (mopoe_mrssm/core.py:83,163)
    assert torch.equal(post[:, 3].feature, post.feature[:, 3])
    again = cat_states([post[:, :4], post[:, 4:]], dim=1)
    assert torch.equal(again.feature, post.feature)
    # fused KL == KL recomputed from the returned probabilities
    fused = kl_divergence(q=post.distribution.independent(1), p=prior.distribution.independent(1), use_balancing=True)
    sliced = kl_divergence(q=post[:, :].distribution.independent(1), p=prior[:, :].distribution.independent(1), use_balancing=True)
    torch.testing.assert_close(fused, sliced, rtol=1e-5, atol=1e-7)
    with torch.no_grad():
        imag = model.rollout_transition(actions=batch[0][:, :4], prev_state=post[:, -1])
    assert imag.feature.shape == (B, 4, 48)
    with pytest.raises(RuntimeError, match="forward-only"):
        model.rollout_transition(actions=batch[0][:, :4].requires_grad_(True), prev_state=post[:, -1])


def test_autocast_selects_the_bf16_path_and_training_reduces_the_loss():
    from multimodal_mtrssm_b200 import _lib

    torch.manual_seed(0)
    model = H.build_mtrssm_model().cuda()
    assert model._precision() == _lib.PRECISION_FP32
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert model._precision() == _lib.PRECISION_BF16_FUSED  # MMTRSSM: bf16 path with the fused (tcgen05) backward
    model.rollout_precision = "bf16_two_kernel"
    assert model._precision() == _lib.PRECISION_BF16
    model.rollout_precision = "bf16"
    assert model._precision() == _lib.PRECISION_BF16_FUSED
    B, T = 32, 12
    g = torch.Generator().manual_seed(1)
    obs = torch.rand(B, T, 1, 32, 32, generator=g).cuda() * 2 - 1
    batch = (torch.randn(B, T, 6, generator=g).cuda(), obs, obs.flip(-1), None, obs, obs.flip(-1))
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    losses = []
    for _ in range(25):
        opt.zero_grad(set_to_none=True)
        out = model.training_step(batch, 0)
        out["loss"].backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        opt.step()
        losses.append(float(out["loss"]))
    assert losses[-1] < losses[0] and all(map(lambda v: v == v, losses)), losses
    assert set(out) == {"loss", "train/loss", "train/recon", "train/recon/audio", "train/recon/vision", "train/kl", "train/kl_h"}


def test_wide_mrssm_model_trains_under_autocast_and_fails_loudly_in_fp32():
    """MoPoE_MRSSM with deterministic_size = hidden_size = 512 (BASELINE.json cfg3) behind the reference's class interface: the
    fused wide kernels are the bf16 tensor-core path, selected under autocast (the reference trains with 16-mixed); the
    fp32-parity policy is not built for this size and must raise, not fall back."""
    torch.manual_seed(0)
    model = H.build_mrssm_model(512).cuda()
    B, T = 24, 6
    g = torch.Generator().manual_seed(1)
    obs = torch.rand(B, T, 1, 32, 32, generator=g).cuda() * 2 - 1
    batch = (torch.randn(B, T, 6, generator=g).cuda(), obs, obs.flip(-1), None, obs, obs.flip(-1))
    with pytest.raises(RuntimeError, match="RSSM_PRECISION_BF16 only"):
        model.training_step(batch, 0)
    opt = torch.optim.AdamW(model.parameters(), lr=3e-4)
    losses = []
    with torch.autocast("cuda", dtype=torch.bfloat16):
        for _ in range(15):
            opt.zero_grad(set_to_none=True)
            out = model.training_step(batch, 0)
            out["loss"].backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
            opt.step()
            losses.append(float(out["loss"].detach()))
        with torch.no_grad():
            post, _ = model.rollout_representation(actions=batch[0], observations=(obs, obs.flip(-1)),
                                                   prev_state=model.initial_state((obs[:, 0], obs.flip(-1)[:, 0])))
            imag = model.rollout_transition(actions=batch[0][:, :3], prev_state=post[:, -1])
    assert post.feature.shape == (B, T, 528) and imag.feature.shape == (B, 3, 528)
    assert losses[-1] < losses[0] and all(map(lambda v: v == v, losses)), losses
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for n, p in model.named_parameters()
               if not n.startswith("representation."))


def test_cuda_graph_training_step_tracks_the_eager_step():
    """dp.GraphedTrainStep: the whole training step (encoders, initial state, fused rollout, decoders, fused likelihood, backward,
    clip, AdamW) replayed as ONE CUDA graph.  With the noise-free parts equal, graph and eager runs from the same initial weights
    on the same batch follow the same loss curve (they draw different categorical noise: Philox offsets differ under capture), the
    graph's inputs are really inputs (a different batch changes the loss), and wrong shapes / non-capturable optimisers raise."""
    import copy

    from multimodal_mtrssm_b200 import dp

    torch.manual_seed(0)
    eager = H.build_mtrssm_model().cuda()
    graphed = copy.deepcopy(eager)
    B, T = 16, 10
    g = torch.Generator().manual_seed(1)
    obs = torch.rand(B, T, 1, 32, 32, generator=g).cuda() * 2 - 1
    act = torch.randn(B, T, 6, generator=g).cuda()
    batch = (act, obs, obs.flip(-1), act.clone(), obs, obs.flip(-1))
    opt_e = torch.optim.AdamW(eager.parameters(), lr=1e-3)
    opt_g = torch.optim.AdamW(graphed.parameters(), lr=1e-3, capturable=True)
    with pytest.raises(RuntimeError, match="capturable"):
        dp.GraphedTrainStep(graphed, batch, torch.optim.AdamW(graphed.parameters(), lr=1e-3), dp.FlatGradBucket(graphed.parameters()))
    w0 = copy.deepcopy(graphed.state_dict())
    step = dp.GraphedTrainStep(graphed, batch, opt_g, dp.FlatGradBucket(graphed.parameters()), warmup=2)
    # the warm-up and the capture trained `graphed` for 2 steps (capture itself executes nothing): restart both from w0
    graphed.load_state_dict(w0)
    eager.load_state_dict(w0)
    bucket_e = dp.FlatGradBucket(eager.parameters())
    le, lg = [], []
    for _ in range(30):
        le.append(float(dp.train_step(eager, batch, opt_e, bucket_e)["loss"]))
        lg.append(float(step(batch)["loss"]))
    assert all(v == v for v in lg) and lg[-1] < lg[0] - 1.0, lg
    assert abs(lg[0] - le[0]) < 0.02 * abs(le[0]), (lg[0], le[0])       # same weights, same batch; only the draws differ
    assert abs(lg[-1] - le[-1]) < 0.05 * abs(le[0] - le[-1]) + 0.02 * abs(le[-1]), (lg[-1], le[-1])
    other = tuple(t.clone() for t in batch)
    other[1].mul_(-1.0)
    other[4].mul_(-1.0)
    assert abs(float(step(other)["train/recon/audio"]) - float(step(batch)["train/recon/audio"])) > 1e-3
    with pytest.raises(ValueError, match="shape"):
        step(tuple(t[:, :5] for t in batch))


def test_pinned_prefetcher_delivers_every_batch_in_order():
    from multimodal_mtrssm_b200 import dp

    dev = torch.device("cuda", 0)
    hosts = [{"x": torch.full((1 << 20,), float(i)).pin_memory(), "y": torch.arange(8, dtype=torch.float32).add_(i).pin_memory()} for i in range(5)]
    pre = dp.PinnedPrefetcher(hosts[0], dev)
    pre.submit(hosts[0])
    sums = []
    for i in range(5):
        slot, d = pre.next()
        if i + 1 < 5:
            pre.submit(hosts[i + 1])
        sums.append((d["x"].mean() + d["y"][0]).reshape(1))     # consumer work on the compute stream
        pre.release(slot)
    assert torch.cat(sums).cpu().tolist() == [2.0 * i for i in range(5)]
    with pytest.raises(RuntimeError, match="nothing submitted"):
        pre.next()
    with pytest.raises(RuntimeError, match="not pinned"):
        pre.submit({"x": torch.zeros(1 << 20), "y": torch.zeros(8)})
    pre2 = dp.PinnedPrefetcher(hosts[0], dev)
    pre2.submit(hosts[0])
    pre2.submit(hosts[1])
    with pytest.raises(RuntimeError, match="in flight"):
        pre2.submit(hosts[2])


def test_base_rssm_unimodal_rollout_uses_the_fused_kernel_and_matches_the_module_loop():
    """BaseRSSM.rollout_representation (core.py:137-168).  On CUDA it is one fused launch (unimodal flag); its
    deterministic outputs must match the reference-structured per-step loop of the same modules (Transition.forward /
    Representation.forward) teacher-forced on the kernel's draws, and training through it must work."""
    from multimodal_mtrssm_b200 import _lib
    from multimodal_mtrssm_b200.core import BaseRSSM
    from multimodal_mtrssm_b200.state import State

    torch.manual_seed(0)
    m = H.build_mrssm_model().cuda()
    B, T = 24, 7
    g = torch.Generator().manual_seed(3)
    emb = torch.randn(B, T, 64, generator=g).cuda()
    act = torch.randn(B, T, 6, generator=g).cuda()
    m.encode_observation = lambda obs: obs          # unimodal: observations are already embeddings here
    with torch.no_grad():
        h0 = torch.randn(B, 32, generator=g).cuda()
        prev = State(deter=h0, distribution=m.representation.distribution_factory(m.transition.rnn_to_prior_projector(h0)))
    n0 = _lib.launch_count()
    post, prior = BaseRSSM.rollout_representation(m, actions=act, observations=emb, prev_state=prev)
    assert _lib.launch_count() == n0 + 1 and post.feature.shape == (B, T, 48) and prior.stoch.shape == (B, T, 16)
    # per-step module loop, teacher-forced on the kernel's posterior draws
    with torch.no_grad():
        state, deters, probs = prev, [], []
        for t in range(T):
            pr = m.transition(act[:, t], state)
            hid = m.representation.rnn_to_post_projector(torch.cat([pr.deter, emb[:, t]], -1))
            dist = m.representation.distribution_factory(hid)
            state = State(deter=pr.deter, distribution=dist, stoch=post.stoch[:, t])
            deters.append(pr.deter)
            probs.append(dist.probs)
    torch.testing.assert_close(post.deter, torch.stack(deters, 1), rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(post.distribution.probs, torch.stack(probs, 1), rtol=1e-5, atol=2e-6)
    loss = (post.feature ** 2).mean() + m.kl_coeff * __import__("multimodal_mtrssm_b200.distribution", fromlist=["kl_divergence"]).kl_divergence(
        q=post.distribution.independent(1), p=prior.distribution.independent(1), use_balancing=True)
    loss.backward()
    head = m.representation.rnn_to_post_projector
    assert head[0].weight.grad is not None and float(head[0].weight.grad.abs().max()) > 0
    assert m.transition.rnn_cell.weight_hh.grad is not None


def test_base_rssm_shared_step_matches_the_reference_base_class(golden_dir):
    """SURVEY §8 row a6, model level: a concrete subclass of the PRODUCT's BaseRSSM against the same subclass of the REFERENCE's
    BaseRSSM (tests/golden/make_golden.golden_unimodal ran the reference's own initial_state / rollout_representation /
    shared_step, models/core.py:121-221): same state_dict, batch and noise -> same loss dict and the same gradient on every
    parameter; the rollout is ONE fused launch."""
    from multimodal_mtrssm_b200 import _lib
    from multimodal_mtrssm_b200.core import BaseRSSM
    from multimodal_mtrssm_b200.mlp import MLP
    from multimodal_mtrssm_b200.networks import Representation, Transition
    from multimodal_mtrssm_b200.objective import likelihood

    class Unimodal(BaseRSSM):
        def __init__(self, *, encoder, decoder, **kw):
            super().__init__(**kw)
            self.encoder, self.decoder = encoder, decoder

        def encode_observation(self, observation):
            return self.encoder(observation)

        def decode_state(self, state):
            return {"recon": self.decoder(state.feature)}

        def compute_reconstruction_loss(self, reconstructions, targets):
            return {"recon": likelihood(prediction=reconstructions["recon"], target=targets["recon"], event_ndims=3)}

        def get_observations_from_batch(self, batch):
            return batch[1]

        def get_initial_observation(self, observations):
            return observations[:, 0]

        def get_targets_from_batch(self, batch):
            return {"recon": batch[4]}

    g = torch.load(golden_dir / "rssm_unimodal.pt")
    kw = dict(deterministic_size=32, hidden_size=32, distribution_config=[4, 4], activation_name="ELU")
    model = Unimodal(representation=Representation(obs_embed_size=64, **kw), transition=Transition(action_size=6, **kw),
                     init_proj=MLP(in_features=64, out_features=32, num_cells=200, depth=1), kl_coeff=1, use_kl_balancing=True,
                     encoder=H.LinEncoder(), decoder=H.LinDecoder(48))
    model.load_state_dict(g["full_state_dict"], strict=True)
    model.cuda()
    inp = g["inputs"]
    batch = tuple(t.cuda() for t in g["batch"])
    n0 = _lib.launch_count()
    with H.RandQueue([inp["u_z0"], inp["u_post"], inp["u_prior"]]) as q:
        loss = model.shared_step(batch)
        assert not q.values
    assert _lib.launch_count() == n0 + 2  # the fused rollout + the fused likelihood
    rep = H.Report("BaseRSSM (unimodal) shared_step vs the reference's BaseRSSM.shared_step")
    for k, v in g["loss"].items():
        rep.check(k, loss[k], v, rtol=1e-5, atol=1e-6)
    loss["loss"].backward()
    for k, p in model.named_parameters():
        rep.check("d " + k, p.grad if p.grad is not None else torch.zeros_like(p), g["full_grads"][k], **_tol(g["full_grads"][k]))
    rep.finish()
    # imagination through BaseRSSM.rollout_transition (core.py:170-185): one fused launch as well
    from multimodal_mtrssm_b200.distribution import Distribution
    from multimodal_mtrssm_b200.state import State

    out, im = g["outputs"], g["imagine"]
    prev = State(deter=out["deter"][:, -1].cuda(), stoch=out["post_stoch"][:, -1].cuda(), distribution=Distribution(out["post_probs"][:, -1].cuda()))
    n0 = _lib.launch_count()
    with torch.no_grad(), H.RandQueue([im["u"]]):
        got = BaseRSSM.rollout_transition(model, actions=im["actions"].cuda(), prev_state=prev)
    assert _lib.launch_count() == n0 + 1
    torch.testing.assert_close(got.deter.cpu(), im["deter"], rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(got.stoch.cpu(), im["stoch"], rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(got.distribution.probs.cpu(), im["probs"], rtol=1e-5, atol=2e-6)


def test_mmtrssm_hoisted_obs_projection_trains_like_the_plain_model(golden_dir):
    """SURVEY §8 f2 at model level: `hoist_obs_projection = True` (same parameters, same state_dict) under autocast.  The two paths
    round differently, so a few of the 2880 categorical draws of this batch may flip (the op-level test against the teacher-forced
    oracle is the parity test, tests/test_rollout_gpu.py); here: same loss dict within 2 %, a gradient on exactly the same
    parameters (incl. both halves of the modality heads' first-layer weights and the encoders), each pointing the same way."""
    g = torch.load(golden_dir / "mtrssm_cfg2.pt")
    inp = g["inputs"]
    batch = tuple(t.cuda() for t in H.golden_batch(g))
    noise = [inp["u_h0"], inp["u_l0"], inp["u_post_l"], inp["u_post_h"], inp["u_prior_l"], inp["u_prior_h"]]
    results = {}
    for hoist in (False, True):
        model = H.build_mtrssm_model()
        model.load_state_dict(g["full_state_dict"], strict=True)
        model.cuda()
        model.hoist_obs_projection = hoist
        with H.RandQueue(noise) as q, torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model.shared_step(batch)
            assert not q.values
        loss["loss"].backward()
        results[hoist] = (loss, {k: p.grad for k, p in model.named_parameters() if p.grad is not None})
    rep = H.Report("MoPoE_MMTRSSM hoist_obs_projection vs plain bf16 path")
    for k, v in results[False][0].items():
        rep.check(k, results[True][0][k], v, rtol=2e-2, atol=2e-2)
    rep.finish()
    assert set(results[True][1]) == set(results[False][1])
    for k, gref in results[False][1].items():
        got = results[True][1][k]
        assert bool(torch.isfinite(got).all()), k
        cos = torch.nn.functional.cosine_similarity(got.flatten().float(), gref.flatten().float(), dim=0)
        assert float(cos) > 0.95, (k, float(cos))
