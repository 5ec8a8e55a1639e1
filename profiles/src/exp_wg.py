import sys, torch
sys.path.insert(0, '.')
from bench import DirectMtrssm
from multimodal_mtrssm_b200 import _lib
run = DirectMtrssm(37888, 30, _lib.PRECISION_BF16, torch.device('cuda'))
run.fwd(); run.bwd_data()
for _ in range(6): run.wgrad()
torch.cuda.synchronize()
print("ok")
