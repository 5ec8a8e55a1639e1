import sys, torch, ctypes as C
sys.path.insert(0, '.')
from bench import DirectMtrssm
from multimodal_mtrssm_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
run = DirectMtrssm(B, 30, _lib.PRECISION_BF16, torch.device('cuda'))
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
print("fwd (train, saved)   ", timeit(run.fwd))
saved_ptr = run.c_out.saved
run.c_out.saved = None
print("fwd (no saved record)", timeit(run.fwd))
run.c_out.saved = saved_ptr
print("bwd_data             ", timeit(run.bwd_data))
print("wgrad                ", timeit(run.wgrad))
