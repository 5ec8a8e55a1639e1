import sys, traceback
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributions as td
from tests import helpers as H
from multimodal_mtrssm_b200.distribution import kl_divergence, Distribution

torch.manual_seed(0)
model = H.build_mtrssm_model().cuda()
B, T = 16, 10
g = torch.Generator().manual_seed(1)
obs = torch.rand(B, T, 1, 32, 32, generator=g).cuda() * 2 - 1
act = torch.randn(B, T, 6, generator=g).cuda()
with torch.no_grad():
    st0 = model.initial_state(model.get_initial_observation((obs, obs.flip(-1))))
probs = torch.softmax(torch.randn(B, T, 4, 4, device="cuda"), -1)

def a_cat():
    return td.OneHotCategoricalStraightThrough(probs=probs, validate_args=False).probs
def a_ind():
    return Distribution(probs).independent(1).base_dist.probs
def a_roll_only():
    po, pr = model.rollout_representation(actions=act, observations=(obs, obs.flip(-1)), prev_state=st0)
    return po.feature
def a_roll_ind():
    po, pr = model.rollout_representation(actions=act, observations=(obs, obs.flip(-1)), prev_state=st0)
    q = po.distribution_l.independent(1)
    return po.feature
def a_kl():
    po, pr = model.rollout_representation(actions=act, observations=(obs, obs.flip(-1)), prev_state=st0)
    return kl_divergence(q=po.distribution_l.independent(1), p=pr.distribution_l.independent(1), use_balancing=True)

for name, fn in [("cat", a_cat), ("ind", a_ind), ("roll_only", a_roll_only), ("roll_ind", a_roll_ind), ("kl", a_kl)]:
    try:
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2): fn()
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            out = fn()
        gr.replay(); torch.cuda.synchronize()
        print(name, "OK", flush=True)
    except Exception:  # noqa: BLE001
        print(name, "FAILED", flush=True)
        traceback.print_exc()
        break
