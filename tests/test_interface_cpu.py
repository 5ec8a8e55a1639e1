"""CPU tests of the host-side mirror of the reference interface, the C-ABI surface and the
data-parallel plumbing.  No GPU compute is launched here."""

import os
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch
import torch.distributions as td

from oracle import rssm_oracle as O
from tests import helpers as H

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "multimodal_mtrssm_b200"


# ---- containers (models/state.py, models/mmtrssm/state.py) ---------------------------------------------------------
def _state(B=3, T=None):
    from multimodal_mtrssm_b200.distribution import Distribution
    from multimodal_mtrssm_b200.state import State

    shape = (B,) if T is None else (B, T)
    probs = torch.softmax(torch.randn(*shape, 4, 4), -1)
    return State(deter=torch.randn(*shape, 32), distribution=Distribution(probs))


def test_state_samples_in_constructor_and_builds_feature():
    s = _state()
    onehot = s.stoch.detach().reshape(3, 4, 4)
    assert torch.allclose(onehot.sum(-1), torch.ones(3, 4)) and bool(((onehot.round() == 0) | (onehot.round() == 1)).all())
    assert torch.equal(s.feature, torch.cat([s.deter, s.stoch], -1))  # state.py:18


def test_stack_index_cat_roundtrip():
    from multimodal_mtrssm_b200.state import cat_states, stack_states

    steps = [_state() for _ in range(5)]
    st = stack_states(steps, dim=1)
    assert st.deter.shape == (3, 5, 32) and st.distribution.probs.shape == (3, 5, 4, 4)
    for t, s in enumerate(steps):
        assert torch.equal(st[:, t].deter, s.deter) and torch.equal(st[:, t].stoch, s.stoch)
        assert torch.equal(st[:, t].distribution.probs, s.distribution.probs)
    both = cat_states([st[:, :2], st[:, 2:]], dim=1)
    assert torch.equal(both.feature, st.feature)
    assert len(list(iter(st))) == 3 and st.unsqueeze(0).squeeze(0).deter.shape == st.deter.shape
    assert st.detach().deter.requires_grad is False and st.clone().deter.data_ptr() != st.deter.data_ptr()


def test_mtstate_feature_order_and_cat_takes_last_hidden():
    from multimodal_mtrssm_b200.distribution import Distribution
    from multimodal_mtrssm_b200.mtstate import MTState, cat_mtstates, stack_mtstates

    def mk():
        return MTState(
            deter_h=torch.randn(2, 32), deter_l=torch.randn(2, 32), hidden_h=torch.randn(2, 32), hidden_l=torch.randn(2, 32),
            distribution_h=Distribution(torch.softmax(torch.randn(2, 8, 2), -1)), distribution_l=Distribution(torch.softmax(torch.randn(2, 4, 4), -1)),
        )

    s = mk()
    assert torch.equal(s.feature, torch.cat([s.deter_h, s.stoch_h, s.deter_l, s.stoch_l], -1))  # mmtrssm/state.py:51
    st = stack_mtstates([mk() for _ in range(4)], dim=1)
    assert st.feature.shape == (2, 4, 96) and st.hidden_l.shape == (2, 4, 32)
    a, b = st[:, :1], st[:, 1:]
    cat = cat_mtstates([a, b], dim=1)
    assert torch.equal(cat.feature, st.feature) and torch.equal(cat.hidden_h, b.hidden_h)  # state.py:237-238
    assert torch.equal(st.clone().distribution_h.probs, st.distribution_h.probs)


# ---- third-party stand-ins (A1..A6) --------------------------------------------------------------------------------
def test_distribution_shim_matches_torch_distributions():
    from multimodal_mtrssm_b200.distribution import Distribution, MultiOneHotFactory, kl_divergence

    torch.manual_seed(0)
    logits_q, logits_p = torch.randn(6, 5, 16, requires_grad=True), torch.randn(6, 5, 16, requires_grad=True)
    fac = MultiOneHotFactory(class_size=2, category_size=8)
    q, p = fac(logits_q), fac(logits_p)
    assert q.probs.shape == (6, 5, 8, 2)  # A1: [category, class], softmax over class
    want = td.kl_divergence(q.independent(1), p.independent(1)).mean()
    torch.testing.assert_close(kl_divergence(q=q.independent(1), p=p.independent(1), use_balancing=False), want)
    bal = kl_divergence(q=q.independent(1), p=p.independent(1), use_balancing=True)
    torch.testing.assert_close(bal, O.kl_per_sample(q.probs, p.probs, True).mean())
    gq, gp = torch.autograd.grad(bal, [logits_q, logits_p], retain_graph=True)
    gq1, gp1 = torch.autograd.grad(O.kl_per_sample(q.probs, p.probs, False).mean(), [logits_q, logits_p], retain_graph=True)
    torch.testing.assert_close(gq, 0.2 * gq1)  # A5 balancing split
    torch.testing.assert_close(gp, 0.8 * gp1)
    z = q.rsample()
    assert z.shape == (6, 5, 16)  # A2: flattened
    (gz,) = torch.autograd.grad((z * torch.arange(16.0)).sum(), q.probs)
    torch.testing.assert_close(gz, torch.arange(16.0).reshape(8, 2).expand(6, 5, 8, 2))  # straight-through


def test_mlp_stand_in_layout_and_default_activation():
    from multimodal_mtrssm_b200.mlp import MLP

    m = MLP(in_features=64, out_features=32, num_cells=200, depth=1)
    assert list(m.state_dict()) == ["0.weight", "0.bias", "2.weight", "2.bias"] and isinstance(m[1], torch.nn.Tanh)  # A6
    assert isinstance(MLP(4, 2, 8, activation_class="torch.nn.ELU")[1], torch.nn.ELU)


def test_likelihood_matches_reference_formula():
    from multimodal_mtrssm_b200.objective import likelihood

    pred, tgt = torch.randn(2, 3, 1, 8, 8), torch.randn(2, 3, 1, 8, 8)
    want = -td.Independent(td.Normal(pred, 1.0), 3).log_prob(tgt).mean()  # objective.py:21-23
    torch.testing.assert_close(likelihood(pred, tgt, event_ndims=3), want)
    want2 = -td.Independent(td.Normal(pred, 0.5), 3).log_prob(tgt).mean()
    torch.testing.assert_close(likelihood(pred, tgt, event_ndims=3, scale=0.5), want2)


# ---- single-step modules against the oracle ------------------------------------------------------------------------
def test_transition_and_mtrnn_single_step_match_oracle():
    from multimodal_mtrssm_b200.distribution import Distribution
    from multimodal_mtrssm_b200.mopoe_mmtrssm import MTRNN
    from multimodal_mtrssm_b200.networks import Transition
    from multimodal_mtrssm_b200.state import State

    torch.manual_seed(1)
    tr = Transition(deterministic_size=32, hidden_size=32, action_size=6, distribution_config=[4, 4], activation_name="ELU")
    params = {"transition." + k: v for k, v in tr.state_dict().items()}
    prev = State(deter=torch.randn(5, 32), distribution=Distribution(torch.softmax(torch.randn(5, 4, 4), -1)))
    act = torch.randn(5, 6)
    got = tr(act, prev)
    deter, probs, _ = O.mrssm_transition(params, act, prev.deter, prev.stoch, 4, 4)
    torch.testing.assert_close(got.deter, deter)
    torch.testing.assert_close(got.distribution.probs, probs)
    cell = MTRNN(input_dim=38, hidden_dim=32, tau=2.0)
    cp = {"l_rnn." + k: v for k, v in cell.state_dict().items()}
    x, d, u = torch.randn(5, 38), torch.randn(5, 32), torch.randn(5, 32)
    cell.hidden = u
    want_d, want_u = O.mtrnn(cp, "l_rnn", x, d, u, 2.0)
    torch.testing.assert_close(cell(x, d), want_d)
    torch.testing.assert_close(cell.hidden, want_u)


# ---- error behaviour (SURVEY.md §8(b) "Errors") ----------------------------------------------------------------------
def test_reference_error_behaviour():
    from multimodal_mtrssm_b200.mopoe_mmtrssm import MTRNN
    from multimodal_mtrssm_b200.networks import Representation

    with pytest.raises(ValueError, match="2 elements"):
        Representation(deterministic_size=32, hidden_size=32, obs_embed_size=64, distribution_config=[4, 4, 4])
    with pytest.raises(AssertionError, match="tau"):
        MTRNN(4, 4, tau=1.0)
    for model in (H.build_mrssm_model(), H.build_mtrssm_model()):
        with pytest.raises(TypeError, match="requires tuple"):
            model.rollout_representation(actions=torch.zeros(1, 2, 6), observations=torch.zeros(1, 2, 1, 32, 32), prev_state=None)


def test_no_cpu_fallback():
    """The product path is CUDA-only: CPU tensors raise instead of silently running an eager/oracle path."""
    from multimodal_mtrssm_b200 import params as P
    from multimodal_mtrssm_b200 import rollout_ops as R

    with pytest.raises(RuntimeError):
        R.mrssm_rollout(P.mrssm_weight_list(H.make_params(H.MR_SHAPES)), **H.mrssm_inputs(2, 2))
    src = "\n".join(p.read_text() for p in PKG.rglob("*.py")) + "\n".join(p.read_text() for p in (PKG / "csrc").glob("*"))
    assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), "the product must not import the oracle"


# ---- state_dict / class-path compatibility -------------------------------------------------------------------------
def test_state_dict_keys_match_reference(golden_dir):
    for name, build in (("mrssm_default.pt", H.build_mrssm_model), ("mtrssm_default.pt", H.build_mtrssm_model)):
        g = torch.load(golden_dir / name)
        model = build()
        ours = model.state_dict()
        assert set(ours) == set(g["full_state_dict"]), set(ours) ^ set(g["full_state_dict"])
        assert all(ours[k].shape == v.shape for k, v in g["full_state_dict"].items())
        model.load_state_dict(g["full_state_dict"], strict=True)
        # alias kept: representation.* is audio_representation.* (mopoe_mrssm/core.py:49,55)
        assert model.representation is model.audio_representation
    mt = H.build_mtrssm_model().state_dict()
    for dead in ("transition.rnn_cell.weight_ih", "l_posterior.0.weight", "l_rnn._d2h.weight", "h_rnn._input2h.bias"):
        assert dead in mt


def test_class_paths_and_yaml_configs_instantiate():
    from multimodal_mtrssm_b200 import compat

    compat.install()
    import importlib

    for path, cls in (("multimodal_rssm.models.mrssm.mopoe_mrssm", "MoPoE_MRSSM"), ("multimodal_rssm.models.mmtrssm.mopoe_mmtrssm", "MoPoE_MMTRSSM"),
                      ("multimodal_rssm.models.networks", "Transition"), ("multimodal_rssm.models.mmtrssm", "cat_mtstates"),
                      ("multimodal_rssm.models", "stack_states"), ("multimodal_rssm.models.objective", "likelihood")):
        assert hasattr(importlib.import_module(path), cls)
    yamls = [PKG / "configs" / "mopoe_mrssm_default.yaml", PKG / "configs" / "mopoe_mmtrssm_default.yaml"]
    ref = Path("/root/reference/src/multimodal_rssm/models")
    if ref.exists():  # the reference's own configs, where the reference checkout is available
        yamls += [ref / "mrssm/mopoe_mrssm/configs/default.yaml", ref / "mmtrssm/mopoe_mmtrssm/configs/default.yaml"]
    for y in yamls:
        model = compat.load_model(y)
        assert type(model).__module__.startswith("multimodal_mtrssm_b200")
        assert len(model.rollout_weights()) in (20, 28)


# ---- C ABI ----------------------------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    from multimodal_mtrssm_b200 import _lib

    header = (ROOT / "include" / "rssm_rollout.h").read_text()
    declared = set(re.findall(r"\b(rssm_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    handle = _lib.lib()  # builds with nvcc if needed; loads without a GPU
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.rssm_abi_version() == _lib.ABI_VERSION
    import ctypes as C

    assert C.sizeof(_lib.MrssmWeights) == 20 * 8 and C.sizeof(_lib.MtrssmWeights) == 28 * 8
    assert C.sizeof(_lib.NllPair) == 9 * 8  # 2 pointers, 2 size_t, float (+pad), 4 pointers
    assert handle.rssm_gaussian_nll_workspace_bytes() >= 64 + 8 * 148 * 4


def test_gaussian_nll_abi_rejects_bad_arguments_without_a_gpu():
    """Argument checks of the likelihood entry points run on the host before any launch (include/rssm_rollout.h)."""
    import ctypes as C

    from multimodal_mtrssm_b200 import _lib

    handle = _lib.lib()
    pair = (_lib.NllPair * 1)()
    pair[0].prediction, pair[0].target, pair[0].n_elems, pair[0].n_batch, pair[0].scale, pair[0].loss = 256, 512, 10, 3, 1.0, 1024
    assert handle.rssm_gaussian_nll_fwd(pair, 1, 0, C.c_void_p(4096), 1 << 20, None) != 0
    assert b"multiple of n_batch" in handle.rssm_last_error()
    pair[0].n_batch = 5
    assert handle.rssm_gaussian_nll_fwd(pair, 1, 7, C.c_void_p(4096), 1 << 20, None) != 0 and b"pred_dtype" in handle.rssm_last_error()
    assert handle.rssm_gaussian_nll_fwd(pair, 1, 0, C.c_void_p(4096), 8, None) != 0 and b"workspace" in handle.rssm_last_error()
    assert handle.rssm_gaussian_nll_fwd(pair, 9, 0, C.c_void_p(4096), 1 << 20, None) != 0 and b"n_pairs" in handle.rssm_last_error()
    pair[0].prediction = 260
    assert handle.rssm_gaussian_nll_fwd(pair, 1, 0, C.c_void_p(4096), 1 << 20, None) != 0 and b"aligned" in handle.rssm_last_error()
    pair[0].prediction, pair[0].d_prediction = 256, None
    assert handle.rssm_gaussian_nll_bwd(pair, 1, 0, None) != 0 and b"d_prediction" in handle.rssm_last_error()


def test_oracle_likelihood_is_the_closed_form_the_kernel_computes():
    from multimodal_mtrssm_b200.objective import likelihood
    from oracle import rssm_oracle as O

    g = torch.Generator().manual_seed(0)
    pred, tgt = torch.randn(3, 4, 1, 8, 8, generator=g), torch.randn(3, 4, 1, 8, 8, generator=g)
    for scale in (1.0, 0.3):
        torch.testing.assert_close(likelihood(pred, tgt, 3, scale), O.likelihood(pred, tgt, 3, scale), rtol=1e-6, atol=0)


def test_wide_family_abi_size_queries_and_struct_layout():
    """ABI v2 (wide MoPoE-MRSSM sizes): record / workspace sizes are host computations that must work without a GPU for the
    record, reject unsupported dims with 0 (the launch reports the error), and the ctypes structs carry the workspace fields."""
    import ctypes as C

    from multimodal_mtrssm_b200 import _lib, rollout_ops

    assert C.sizeof(_lib.MrssmOutputs) == 8 * 8 and C.sizeof(_lib.MrssmInputGrads) == 8 * 8  # 7 pointers + size_t
    d512 = _lib.MrssmDims(B=1024, T=64, A=6, E=64, D=512, H=512, C=4, K=4, precision=_lib.PRECISION_BF16)
    planes = 64 * 10 * 8 * 512 * 128 * 2          # [T][plane][block][D/8][128][8] bf16
    logits = 8 * 128 * 64 * 32 * 4                # [blocks*128][T][32] fp32
    assert _lib.mrssm_saved_bytes(d512) == planes + logits
    assert rollout_ops._mr_saved_shape(1024, 64, 512)[0] * 2 == planes + logits
    ragged = _lib.MrssmDims(B=200, T=6, A=6, E=64, D=384, H=384, C=8, K=2, precision=_lib.PRECISION_BF16)
    assert _lib.mrssm_saved_bytes(ragged) == rollout_ops._mr_saved_shape(200, 6, 384)[0] * 2  # padded to 2 blocks of 128
    small = _lib.MrssmDims(B=8, T=30, A=6, E=64, D=32, H=32, C=4, K=4, precision=_lib.PRECISION_FP32)
    assert _lib.mrssm_saved_bytes(small) == 8 * 30 * _lib.MRSSM_SAVED_FLOATS * 4 and _lib.mrssm_workspace_bytes(small, False) == 0
    for bad in (dict(D=96, H=96), dict(D=512, H=256), dict(D=1024, H=1024), dict(D=512, H=512, precision=_lib.PRECISION_FP32),
                dict(D=512, H=512, A=9), dict(D=512, H=512, C=4, K=8)):
        kw = dict(B=8, T=4, A=6, E=64, D=512, H=512, C=4, K=4, precision=_lib.PRECISION_BF16)
        kw.update(bad)
        assert _lib.mrssm_saved_bytes(_lib.MrssmDims(**kw)) == 0, bad


def test_wide_family_fake_shapes_and_size_check():
    """Shape inference of the custom ops (meta tensors, no kernel) follows deterministic_size, and unsupported sizes are
    rejected by the host wrapper before any launch."""
    from multimodal_mtrssm_b200 import rollout_ops

    D, B, T, K = 512, 5, 3, 4
    meta = lambda *s: torch.empty(*s, device="meta")  # noqa: E731
    shapes = H.mr_shapes(D)
    from multimodal_mtrssm_b200.params import MR_STATE_KEYS

    weights = [meta(*shapes[k]) for k in MR_STATE_KEYS]
    out = torch.ops.mtrssm_b200.mrssm_rollout(weights, meta(B, T, 6), meta(B, T, 64), meta(B, T, 64), meta(B, D), meta(B, 16),
                                              meta(B, T, 4), None, K, 1, 0.2, 0.8, True, False)
    assert out[0].shape == (B, T, D + 16) and out[1].shape == (B, T, 4, 4) and out[5].dtype == torch.bfloat16
    assert out[5].numel() * 2 == 3 * 10 * 1 * D * 128 * 2 + 128 * 3 * 32 * 4
    with pytest.raises(RuntimeError, match="supports"):
        rollout_ops._mr_check([torch.empty(s) for s in H.mr_shapes(96).values()], torch.empty(2, 3, 64), torch.empty(2, 96), torch.empty(2, 16))


# ---- data parallel plumbing (gloo, world size 2) ------------------------------------------------------------------------
DP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO"])
from multimodal_mtrssm_b200.dp import FlatGradBucket, broadcast_parameters, reduce_metrics, shard_batch
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["PORT"], rank=rank, world_size=world)
torch.manual_seed(100 + rank)                      # different init per rank on purpose
net = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
dead = torch.nn.Linear(3, 3)                        # never used: grad stays None (MMTRSSM's dummy transition)
broadcast_parameters(net)
g = torch.Generator().manual_seed(5)
x, y = torch.randn(8, 6, generator=g), torch.randn(8, 1, generator=g)
xs, ys = shard_batch((x, y), rank, world)
loss = (net(xs) - ys).square().mean()
loss.backward()
bucket = FlatGradBucket(list(net.parameters()) + list(dead.parameters()))
n = bucket.allreduce()
ref = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
ref.load_state_dict(net.state_dict())
(ref(x) - y).square().mean().backward()            # full-batch gradient
for a, b in zip(net.parameters(), ref.parameters()):
    torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-6)
assert all(p.grad is None for p in dead.parameters()) and n == sum(p.numel() for p in net.parameters())
m = reduce_metrics({"loss": loss, "k": torch.tensor(float(rank))})
torch.testing.assert_close(m["k"], torch.tensor(0.5))
torch.testing.assert_close(m["loss"], (ref(x) - y).square().mean().detach(), rtol=1e-5, atol=1e-6)
# readiness-ordered overlapped buckets: three steps (recording step + two hook-driven ones), tiny buckets -> several collectives
from multimodal_mtrssm_b200.dp import OverlappedGradBuckets, train_step
class Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.enc, self.mid, self.dec, self.dead = torch.nn.Linear(6, 8), torch.nn.Linear(8, 8), torch.nn.Linear(8, 1), torch.nn.Linear(3, 3)
    def training_step(self, batch, _):
        x, y = batch
        return {"loss": (self.dec(torch.tanh(self.mid(torch.tanh(self.enc(x))))) - y).square().mean()}
torch.manual_seed(7 + rank)
toy, full = Toy(), Toy()
broadcast_parameters(toy)
full.load_state_dict(toy.state_dict())
ob = OverlappedGradBuckets(toy.parameters(), bucket_bytes=64)
opt_t, opt_f = torch.optim.SGD(toy.parameters(), lr=0.1), torch.optim.SGD(full.parameters(), lr=0.1)
for step in range(3):
    train_step(toy, (xs, ys), opt_t, ob, clip=None)
    opt_f.zero_grad()
    full.training_step((x, y), 0)["loss"].backward()
    for (name, a), b in zip(toy.named_parameters(), full.parameters()):
        if name.startswith("dead"):
            assert a.grad is None
        else:
            torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-6)
    opt_f.step()
assert len(ob._buckets) >= 3 and [toy.dec.weight is ob.params[i] for i in ob._buckets[0]].count(True) == 1   # decoder first
assert sum(len(b) for b in ob._buckets) == 6
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_flat_bucket_allreduce_world_size_2(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(DP_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [
        subprocess.Popen([sys.executable, str(script)], env={**os.environ, "RANK": str(r), "WORLD_SIZE": "2", "PORT": port, "REPO": str(ROOT)},
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        for r in range(2)
    ]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


def test_shard_batch_rejects_ragged():
    from multimodal_mtrssm_b200.dp import shard_batch

    with pytest.raises(ValueError, match="divisible"):
        shard_batch((torch.zeros(5, 2),), 0, 2)


def test_oracle_unimodal_switch_follows_the_reference_structured_module_loop():
    """SURVEY 8 row a6 (core.py:137-168): the oracle's `unimodal=True` rollout against the per-step loop of the host modules
    (Transition.forward networks.py:151-173, Representation.forward networks.py:70-84) with the oracle's draws forced."""
    from multimodal_mtrssm_b200 import _lib
    from multimodal_mtrssm_b200.state import State
    from oracle import rssm_oracle as O
    from tests import helpers as H

    import ctypes as C

    assert C.sizeof(_lib.MrssmDims) == 10 * 4  # ... precision, unimodal
    torch.manual_seed(0)
    m = H.build_mrssm_model()
    B, T = 5, 6
    inp = H.mrssm_inputs(B, T, 4, 4)
    params = {k: v.detach() for k, v in m.state_dict().items()}
    res = O.mrssm_rollout(params, C=4, K=4, unimodal=True, **inp)
    with torch.no_grad():
        state = State(deter=inp["h0"], distribution=m.representation.distribution_factory(torch.zeros(B, 16)), stoch=inp["z0"])
        for t in range(T):
            prior = m.transition(inp["actions"][:, t], state)
            dist = m.representation.distribution_factory(
                m.representation.rnn_to_post_projector(torch.cat([prior.deter, inp["embed_a"][:, t]], -1)))
            torch.testing.assert_close(prior.deter, res["deter"][:, t], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(dist.probs, res["post_probs"][:, t], rtol=1e-5, atol=1e-6)
            state = State(deter=prior.deter, distribution=dist, stoch=res["post_stoch"][:, t])
    # the vision inputs / head play no role
    other = dict(inp, embed_v=torch.randn_like(inp["embed_v"]))
    assert torch.equal(O.mrssm_rollout(params, C=4, K=4, unimodal=True, **other)["post_probs"], res["post_probs"])


def test_unimodal_flag_is_rejected_for_the_wide_family_on_the_host():
    """dims.unimodal (BaseRSSM.rollout_representation) is built for the default sizes; the wide family must refuse it before any
    launch (size queries return 0, the launch would report why) rather than silently running the multimodal kernel."""
    from multimodal_mtrssm_b200 import _lib

    ok = _lib.MrssmDims(B=256, T=4, A=6, E=64, D=512, H=512, C=4, K=4, precision=_lib.PRECISION_BF16, unimodal=0)
    bad = _lib.MrssmDims(B=256, T=4, A=6, E=64, D=512, H=512, C=4, K=4, precision=_lib.PRECISION_BF16, unimodal=1)
    assert _lib.mrssm_saved_bytes(ok) > 0 and _lib.mrssm_saved_bytes(bad) == 0
    small = _lib.MrssmDims(B=8, T=4, A=6, E=64, D=32, H=32, C=4, K=4, precision=_lib.PRECISION_FP32, unimodal=1)
    assert _lib.mrssm_saved_bytes(small) == 8 * 4 * _lib.MRSSM_SAVED_FLOATS * 4
