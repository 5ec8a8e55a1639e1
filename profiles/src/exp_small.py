import sys, torch
sys.path.insert(0, '.')
from bench import DirectMtrssm, time_direct
from multimodal_mtrssm_b200 import _lib
for B, T in ((8, 30), (256, 30), (4096, 30), (16384, 30), (256, 512), (37888, 30)):
    row = []
    for prec in (_lib.PRECISION_BF16, _lib.PRECISION_BF16_FUSED):
        run = DirectMtrssm(B, T, prec, torch.device('cuda'))
        r = time_direct(run, 20, 5, 1)
        row.append("%.3f ms (fwd %.3f bwd %.3f wg %.3f)" % (r["total_ms"] / 20, r["fwd_ms"], r["bwd_ms"], r["wgrad_ms"]))
        del run
    print(f"B={B} T={T}: classic {row[0]} | fused {row[1]}")
