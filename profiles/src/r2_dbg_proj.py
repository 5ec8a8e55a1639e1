import sys, torch
sys.path.insert(0, ".")
from multimodal_mtrssm_b200 import _lib, params as P, rollout_ops as R
from tests import helpers as H
from tests.test_rollout_gpu import oracle_mtrssm, mtrssm_upstream, cuda, MT_GRAD_IN
B, T, dims = 48, 8, H.MT_DIMS
params = H.make_params(H.MT_SHAPES)
inp = H.mtrssm_inputs(B, T, dims)
inp["u_prior_l"] = inp["u_prior_h"] = None
up = {k: v for k, v in mtrssm_upstream(B, T, dims).items() if not k.startswith("prior_stoch")}
w = {k: v.cuda().requires_grad_(True) for k, v in params.items()}
x = cuda(inp)
Wa = w["audio_representation.rnn_to_post_projector.0.weight"]
Pa = R.obs_projection(x["embed_a"], Wa).detach().requires_grad_(True)
Pv = R.obs_projection(x["embed_v"], w["vision_representation.rnn_to_post_projector.0.weight"]).detach().requires_grad_(True)
xin = dict(x); xin["embed_a"], xin["embed_v"] = Pa, Pv
got = R.mtrssm_rollout(P.mtrssm_weight_list(w), precision=_lib.PRECISION_BF16_FUSED, obs_projected=True, **xin)
sum((got[k] * up[k].cuda()).sum() for k in up).backward()
# plain path on the same data
w2 = {k: v.cuda().requires_grad_(True) for k, v in params.items()}
x2 = cuda(inp); x2["embed_a"].requires_grad_(True)
got2 = R.mtrssm_rollout(P.mtrssm_weight_list(w2), precision=_lib.PRECISION_BF16_FUSED, **x2)
sum((got2[k] * up[k].cuda()).sum() for k in up).backward()
de_from_dP = Pa.grad @ Wa[:, 32:].detach()
print("same draws:", bool((got["feature"][..., 80:] == got2["feature"][..., 80:]).all()))
print("dP stats", float(Pa.grad.abs().max()), float(Pa.grad.abs().mean()))
print("d e (plain) max", float(x2["embed_a"].grad.abs().max()), " d e from dP max", float(de_from_dP.abs().max()))
err = (de_from_dP - x2["embed_a"].grad).abs().amax(-1)
print("err per t (max over b):", [round(float(v), 4) for v in err.amax(0)])
print("err per tile (b // 16):", [round(float(err[i * 16:(i + 1) * 16].max()), 4) for i in range(3)])
# least-squares dP from plain path: dP_ref = d e . pinv(W1e)
dP_ref = x2["embed_a"].grad @ torch.linalg.pinv(Wa[:, 32:].detach())
print("dP vs ref, row 0 t 0:", Pa.grad[0, 0, :8].tolist(), dP_ref[0, 0, :8].tolist())
print("ratio hist:", torch.quantile((Pa.grad / (dP_ref + 1e-9)).flatten()[:100000], torch.tensor([0.1, 0.5, 0.9], device="cuda")).tolist())
