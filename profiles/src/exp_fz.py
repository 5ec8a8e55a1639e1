import sys, torch
sys.path.insert(0, '.')
from bench import DirectMtrssm
from multimodal_mtrssm_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
run = DirectMtrssm(B, 30, _lib.PRECISION_BF16, torch.device('cuda'))
print("classic: fwd %.3f" % timeit(run.fwd), "bwd_data %.3f" % timeit(run.bwd_data), "wgrad %.3f" % timeit(run.wgrad), end="  |  ")
del run
run = DirectMtrssm(B, 30, _lib.PRECISION_BF16_FUSED, torch.device('cuda'))
print("fused: fwd %.3f" % timeit(run.fwd), "bwd_fused %.3f" % timeit(run.bwd_fused))
