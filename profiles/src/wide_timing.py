"""Phase timestamps of the wide forward / backward kernels (CTA 0, globaltimer ns) at cfg3."""
import os, sys, torch
import numpy as np
os.environ["RSSM_WIDE_TIMING"] = "1"
sys.path.insert(0, ".")
from multimodal_mtrssm_b200 import params as P, rollout_ops as R
from tests import helpers as H
D, B, T, K = 512, 1024, 64, 4
params = {k: v.cuda().requires_grad_(True) for k, v in H.make_params(H.mr_shapes(D)).items()}
inp = {k: v.cuda() for k, v in H.mrssm_inputs(B, T, 4, K, D=D).items()}
inp["u_prior"] = None
w = P.mrssm_weight_list(params)
up = torch.randn(B, T, D + 16, device="cuda")
for _ in range(3):
    out = R.mrssm_rollout(w, class_size=K, precision=1, **inp)
    ((out["feature"] * up).sum() + out["kl"].sum()).backward()
torch.cuda.synchronize()

def report(ws, names, label):
    raw = ws.view(torch.uint8)[-4096:].cpu().numpy().view("uint64")
    ts = raw[raw > 0].astype("int64")
    d = ts[1:] - ts[:-1]
    k = len(names)
    dd = d[: (len(d) // k) * k].reshape(-1, k)
    print(f"{label}: steps timed {dd.shape[0]} (exp={os.environ.get('RSSM_WIDE_EXP', '0')})")
    for i, nm in enumerate(names):
        print(f"  {nm:12s} median {np.median(dd[2:, i]) / 1e3:7.2f} us")
    print(f"  step total median {np.median(dd[2:].sum(1)) / 1e3:.2f} us")

report(R._DEBUG_LAST_WORKSPACE[False], ["A compute", "A barrier", "B compute", "B barrier", "C compute", "C barrier", "D sample", "D sync", "D hid1", "D barrier"], "forward")
report(R._DEBUG_LAST_WORKSPACE[True], ["P1 compute", "P1 barrier", "P2 compute", "P2 barrier", "P3 compute", "P3 barrier", "P4 compute", "P4 barrier"], "backward")
