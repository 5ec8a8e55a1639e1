"""Kernel timings of the MMTRSSM rollout at the bench sizes (CUDA events, direct C-ABI calls, no allocator traffic).
   python profiles/src/r2_quick.py [--no-prior] [--ab]     (--ab: grouped 1 KB output rows vs one dense tensor per output)"""
import sys
import torch
sys.path.insert(0, ".")
from bench import DirectMtrssm
from multimodal_mtrssm_b200 import _lib

prior = "--no-prior" not in sys.argv


def timeit(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


layouts = (True, False) if "--ab" in sys.argv else (False,) if "--dense" in sys.argv else (True,)
sizes = ((37888, 30, 20), (16384, 30, 20), (4096, 30, 20), (256, 30, 20), (8, 30, 20), (256, 512, 5), (16, 512, 5))
if "--only-bench" in sys.argv:
    sizes = ((37888, 30, 20 if "--pipeline" in sys.argv else 3),)
if "--ab-sizes" in sys.argv:
    sizes = ((37888, 30, 20), (4096, 30, 20), (256, 512, 5))
for B, T, n in sizes:
  for grouped in layouts:
    run = DirectMtrssm(B, T, _lib.PRECISION_BF16_FUSED, torch.device("cuda"), prior_sample=prior, grouped=grouped)
    f, b = timeit(run.fwd, n), timeit(run.bwd_fused, n)
    if "--pipeline" in sys.argv:  # the two kernels alternating, as in a training step (each also drains the other's dirty L2 lines)
        print(f"  alternating fwd -> bwd: {timeit(lambda: (run.fwd(), run.bwd_fused()), n):.4f} ms per pair", end="  |  ")
    print("grouped rows " if grouped else "dense outputs", end=" ")
    bytes_step = (3120 if prior else 3120 - 256) * B * T
    print(f"B={B:6d} T={T:4d} prior_sample={prior}: fwd {f:.4f} ms  fused bwd {b:.4f} ms  step {f + b:.4f} ms  "
          f"{B * T / (f + b) * 1e3:.3e} steps/s  frac_of_hbm {bytes_step / (f + b) * 1e3 / 1e9 / 6464.9:.3f}  us/step-of-T {(f + b) / T * 1e3:.2f}")
    del run
