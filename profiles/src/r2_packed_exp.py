"""Timing experiment: the forward kernel writing ONE packed row per (b,t) (feature | hidden | probs | prior draws | kl | saved record)
instead of 11 separate tensors.  RSSM_EXP_ROW_PITCH=368 python profiles/src/r2_packed_exp.py   (forward only: the backward is not
taught the layout in this experiment)."""
import os
import sys
import torch
sys.path.insert(0, ".")
from bench import DirectMtrssm
from multimodal_mtrssm_b200 import _lib

P = int(os.environ.get("RSSM_EXP_ROW_PITCH", "0"))
OFF = {"feature": 0, "hidden_h": 96, "hidden_l": 128, "prior_probs_h": 160, "prior_probs_l": 176, "post_probs_h": 192, "post_probs_l": 208,
       "prior_stoch_h": 224, "prior_stoch_l": 240, "kl_l": 256, "kl_h": 257, "saved": 260}


def timeit(fn, n=20):
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


for B in (37888, 16384):
    run = DirectMtrssm(B, 30, _lib.PRECISION_BF16_FUSED, torch.device("cuda"))
    if P:
        packed = torch.zeros(B, 30, P, device="cuda")
        for k, off in OFF.items():
            setattr(run.c_out, k, packed.data_ptr() + 4 * off)
    print(f"B={B} row_pitch={P}: fwd {timeit(run.fwd):.4f} ms")
    if P:  # sanity: the packed outputs hold what the kernel wrote (distributions normalise, one-hots)
        pp = packed[..., 208:224].reshape(B, 30, 4, 4)
        print("   post_probs_l rows sum to", float(pp.sum(-1).min()), float(pp.sum(-1).max()), " z_l one-hot:", bool((packed[..., 80:96].reshape(B, 30, 4, 4).sum(-1) == 1).all()))
    del run
