"""Kernel timings of the fp32-parity policy (3 x bf16 split) of the MMTRSSM rollout at the bench size (CUDA events, direct C-ABI calls).
   python profiles/src/r2_fp32_quick.py [B]"""
import sys
import torch
sys.path.insert(0, ".")
from bench import DirectMtrssm
from multimodal_mtrssm_b200 import _lib


def timeit(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


B = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
run = DirectMtrssm(B, 30, _lib.PRECISION_FP32, torch.device("cuda"))
f, b, w = timeit(run.fwd, n), timeit(run.bwd_data, n), timeit(run.wgrad, n)
print(f"fp32-parity policy B={B} T=30: fwd {f:.3f} ms  bwd {b:.3f} ms  wgrad {w:.3f} ms  step {f + b + w:.3f} ms")
