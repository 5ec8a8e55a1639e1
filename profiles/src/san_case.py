import sys, torch
sys.path.insert(0, '.')
from multimodal_mtrssm_b200 import _lib, params as P, rollout_ops as R
from tests import helpers as H
B, T, dims = 70, 5, H.MT_DIMS
params = H.make_params(H.MT_SHAPES)
inp = H.mtrssm_inputs(B, T, dims)
g = torch.Generator().manual_seed(3)
up = torch.randn(B, T, 96, generator=g).cuda()
for prec in (_lib.PRECISION_BF16_FUSED, _lib.PRECISION_BF16, _lib.PRECISION_FP32):
    w = {k: v.cuda().requires_grad_(True) for k, v in params.items()}
    out = R.mtrssm_rollout(P.mtrssm_weight_list(w), precision=prec, **{k: v.cuda() for k, v in inp.items()})
    ((out["feature"] * up).sum() + out["kl_l"].mean() + out["kl_h"].mean()).backward()
    torch.cuda.synchronize()
    print("precision", prec, "ok", float(w["l_prior.0.weight"].grad.abs().sum()))
