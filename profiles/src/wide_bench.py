"""cfg3 timing of the wide MRSSM rollout (B=1024, T=64, D=512): forward, and forward+backward when available."""
import sys, time, torch
sys.path.insert(0, ".")
from multimodal_mtrssm_b200 import params as P, rollout_ops as R
from tests import helpers as H

D, B, T, K = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 64, 4
grad = len(sys.argv) > 3 and sys.argv[3] == "grad"
params = {k: v.cuda().requires_grad_(grad) for k, v in H.make_params(H.mr_shapes(D)).items()}
inp = {k: v.cuda() for k, v in H.mrssm_inputs(B, T, 4, K, D=D).items()}
inp["u_prior"] = None
w = P.mrssm_weight_list(params)
up = torch.randn(B, T, D + 16, device="cuda")

def step():
    out = R.mrssm_rollout(w, class_size=K, precision=1, **inp)
    if grad:
        ((out["feature"] * up).sum() + out["kl"].sum()).backward()
    return out

for _ in range(3):
    step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
n = 10
ev[0].record()
for _ in range(n):
    step()
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / n
flops = 5445632 * B * T * (3 if grad else 1) * (D / 512) ** 2
print(f"wide D={D} B={B} T={T} grad={grad}: {ms:.3f} ms/step, {B*T/ms*1e3:.3e} latent steps/s, {flops/ms/1e9:.1f} TFLOP/s "
      f"({flops/ms/1e9/1384.6*100:.1f} % of sustained bf16 peak)")
