"""Kernel-time breakdown of one full training step (B=256, T=30, autocast) via torch.profiler."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from torch.profiler import profile, ProfilerActivity
from multimodal_mtrssm_b200 import compat, dp, standins, synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
if len(sys.argv) > 2 and sys.argv[2] == "bench":
    torch.backends.cudnn.benchmark = True
T = 30
device = torch.device("cuda", 0)
model = compat.load_model(ROOT / "multimodal_mtrssm_b200" / "configs" / "mopoe_mmtrssm_default.yaml")
standins.materialize(model, model.feature_dim)
model.to(device).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
bucket = dp.FlatGradBucket(model.parameters())
g = torch.Generator().manual_seed(1234)
frames = lambda: (torch.rand(B, T, 1, 32, 32, generator=g) * 2 - 1).to(device)
act = synthetic.actions(B, T, g).to(device)
batch = (act, frames(), frames(), act.clone(), frames(), frames())
def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        dp.train_step(model, batch, opt, bucket)
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(10): step()
ev1.record(); torch.cuda.synchronize()
print("ms_per_step", ev0.elapsed_time(ev1) / 10)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
