"""Per-phase timelines of the MMTRSSM forward and fused backward kernels (clock64 stamps of one tile; -DFZ_TIMING build).
   RSSM_ROLLOUT_LIB=profiles/src/lib_timing.so RSSM_FZ_TIMING=1 python profiles/src/r2_fz_timing.py"""
import sys
import torch
sys.path.insert(0, ".")
from bench import DirectMtrssm
from multimodal_mtrssm_b200 import _lib

for B, T in ((16, 30), (256, 512), (37888, 30)):
    run = DirectMtrssm(B, T, _lib.PRECISION_BF16_FUSED, torch.device("cuda"))
    for _ in range(2):
        run.fwd()
        run.bwd_fused()
    torch.cuda.synchronize()
    del run
