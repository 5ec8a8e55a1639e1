"""Which part of the training step breaks CUDA-graph capture?"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tests import helpers as H
from multimodal_mtrssm_b200 import objective, dp

torch.manual_seed(0)
model = H.build_mtrssm_model().cuda()
B, T = 16, 10
g = torch.Generator().manual_seed(1)
obs = torch.rand(B, T, 1, 32, 32, generator=g).cuda() * 2 - 1
act = torch.randn(B, T, 6, generator=g).cuda()
batch = (act, obs, obs.flip(-1), act.clone(), obs, obs.flip(-1))
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, capturable=True)
params = [p for p in model.parameters()]

def s_nll():
    return objective.likelihood_pairs([obs, obs], [obs.flip(-1), obs], 3)
def s_enc():
    return model.encode_observation((obs, obs))
def s_init():
    return model.initial_state(model.get_initial_observation((obs, obs.flip(-1))))
def s_fwd():
    return model.training_step(batch, 0)["loss"]
def s_fwd_bwd():
    opt.zero_grad(set_to_none=False)
    l = model.training_step(batch, 0)["loss"]; l.backward(); return l
def s_clip():
    l = s_fwd_bwd(); torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 10.0); return l
def s_opt():
    l = s_clip(); opt.step(); return l
def s_fwd_ac():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return model.training_step(batch, 0)["loss"]
def s_fb_ac():
    opt.zero_grad(set_to_none=False)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        l = model.training_step(batch, 0)["loss"]
    l.backward(); return l

with torch.no_grad():
    st0 = model.initial_state(model.get_initial_observation((obs, obs.flip(-1))))
    post0, prior0 = model.rollout_representation(actions=act, observations=(obs, obs.flip(-1)), prev_state=st0)
def s_roll():
    return model.rollout_representation(actions=act, observations=(obs, obs.flip(-1)), prev_state=st0)[0].feature
def s_dec():
    return model.decode_state(post0)["recon/audio"]
def s_recon():
    return model.compute_reconstruction_loss(model.decode_state(post0), model.get_targets_from_batch(batch))["recon"]
def s_kl():
    from multimodal_mtrssm_b200.distribution import kl_divergence
    po, pr = model.rollout_representation(actions=act, observations=(obs, obs.flip(-1)), prev_state=st0)
    return kl_divergence(q=po.distribution_l.independent(1), p=pr.distribution_l.independent(1), use_balancing=True)
def s_rand():
    return torch.rand(4, 5, 3, device="cuda")
def s_op():
    from multimodal_mtrssm_b200 import rollout_ops
    u = {k: torch.rand(B, T, c, device="cuda") for k, c in (("u_post_l", 4), ("u_post_h", 8), ("u_prior_l", 4), ("u_prior_h", 8))}
    ea, ev = torch.randn(B, T, 64, device="cuda"), torch.randn(B, T, 64, device="cuda")
    return rollout_ops.mtrssm_rollout(model.rollout_weights(), actions=act, embed_a=ea, embed_v=ev, **model._state_inputs(st0), **u,
                                      use_kl_balancing=True, **model._kernel_cfg())["feature"]

for name, fn in [("rand", s_rand), ("op", s_op), ("roll", s_roll), ("dec", s_dec), ("recon", s_recon), ("kl", s_kl), ("nll", s_nll), ("enc", s_enc), ("init", s_init), ("fwd", s_fwd), ("fwd_bwd", s_fwd_bwd), ("clip", s_clip), ("opt", s_opt), ("fwd_autocast", s_fwd_ac), ("fwd_bwd_autocast", s_fb_ac)]:
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
    except Exception as e:  # noqa: BLE001
        print(name, "FAILED:", str(e).splitlines()[0][:160], flush=True)
        break
        try:
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001
            pass
