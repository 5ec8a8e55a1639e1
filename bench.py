#!/usr/bin/env python
"""bench.py -- latent steps/s (B*T, rollout fwd + BPTT bwd) of the fused MoPoE-MMTRSSM rollout on B200.

    python bench.py --gpus N --steps K --warmup W            # ours  (torchrun launches it for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference arm: CPU, host cores

Workload (BASELINE.json configs[1]): MoPoE-MMTRSSM at the default.yaml sizes
(hd=ld=32, hs=ls=16, heads 32, E=64, A=6, tau 2/4, KL balancing), T=30, synthetic encoder embeddings, actions
and noise.  default.yaml's batch of 8 cannot occupy a GPU, so the per-GPU batch is 37888 sequences (= 148 SMs x 256:
two full waves of the kernels' 16-sequence warp tiles; weak scaling: every rank processes its own 37888).  The B=8
latency (`default_batch8`) and B = 256 / 4096 / 16384 and cfg4 (`other_workloads`) are reported beside it.

The timed configuration is the one the model API launches (`MoPoE_MMTRSSM.rollout_representation`): the prior MTState's own
draws are made and written (`u_prior_*` / `prior_stoch_*`; `--no-prior-sample` leaves them out and subtracts their 2 x 128
algorithmic bytes per (b,t)).

A step = one pass of the hot path over one batch: the forward rollout kernel + the fused backward kernel (BPTT and the weight
gradients -- tcgen05 MMAs with TMEM accumulators -- in one launch; `--precision bf16` runs the two-kernel backward, `--precision
fp32` the fp32-parity path) (+ one NCCL allreduce of the flat weight-gradient bucket when N > 1).  `value` has the inputs resident in
HBM; `e2e` goes through the public API (`rollout_ops.mtrssm_rollout` + autograd) with pinned-host inputs copied
in and the loss + weight gradients copied out every step.  One JSON line on stdout (rank 0).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "latent_steps_per_sec"
UNIT = "latent steps/s (B*T, rollout fwd+bwd)"
# fp32 mandatory I/O per (b,t): inputs (A + 2E)*4 = 536 B, MMTRSSM outputs 1024 B; fwd+bwd = 2x
FWD_BYTES_PER_BT = 536 + 1024
STEP_BYTES_PER_BT = 2 * FWD_BYTES_PER_BT
PRIOR_STOCH_BYTES_PER_BT = 128  # (hs + ls) * 4: part of the 1024 B of outputs; not moved under --no-prior-sample


def bytes_per_bt(prior_sample: bool) -> tuple[int, int]:
    """(forward, fwd+bwd) algorithmic HBM bytes per (b,t) of the timed configuration."""
    f = FWD_BYTES_PER_BT - (0 if prior_sample else PRIOR_STOCH_BYTES_PER_BT)
    return f, 2 * f
FWD_FLOPS_PER_BT = 33152


def parse() -> argparse.Namespace:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=37888, help="sequences per GPU (default 148 SMs x 256 = two full waves of 16-sequence warp tiles)")
    ap.add_argument("--seq-len", type=int, default=30)
    ap.add_argument("--precision", default="bf16_fused", choices=["bf16", "bf16_fused", "fp32"],
                    help="bf16_fused: bf16 path, one-kernel backward (default); bf16: bf16 path, BPTT kernel + weight-gradient kernel")
    ap.add_argument("--cpu-batch", type=int, default=256, help="sequences in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the fp32-path / B=8 side measurements")
    ap.add_argument("--no-prior-sample", action="store_true",
                    help="do not draw / write the prior MTState's own samples (the model API always does); their bytes are subtracted")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="wall-clock budget of the whole reference-arm run")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.samples: list[list[str]] = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self) -> None:
        assert self.proc is not None and self.proc.stdout is not None
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) == 6:
                self.samples.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
class DirectMtrssm:
    """Pre-allocated buffers + direct C-ABI calls: exactly the three kernels of the hot path, no allocator traffic."""

    def __init__(self, B: int, T: int, precision: int, device: torch.device, prior_sample: bool = True, obs_projected: bool = False,
                 grouped: bool | None = None) -> None:
        from multimodal_mtrssm_b200 import _lib, synthetic
        from multimodal_mtrssm_b200.params import mtrssm_weight_list

        self.lib, self.B, self.T = _lib, B, T
        self.fused = precision == _lib.PRECISION_BF16_FUSED
        self.params = {k: v.to(device) for k, v in synthetic.mtrssm_params().items()}
        self.weights = mtrssm_weight_list(self.params)
        self.inp = {k: v.to(device) for k, v in synthetic.mtrssm_batch(B, T, prior_noise=prior_sample).items()}
        EW = 64
        if obs_projected:  # the kernels take e @ W1[:, 32:].T ([B,T,32]) computed by one GEMM before the loop
            EW = 32
            for m, key in (("audio", "embed_a"), ("vision", "embed_v")):
                w1 = self.params[f"{m}_representation.rnn_to_post_projector.0.weight"]
                self.inp[key] = (self.inp[key] @ w1[:, 32:].T).contiguous()
        g = torch.Generator().manual_seed(7)
        self.d_feature = torch.randn(B, T, 96, generator=g).to(device)
        self.d_kl = torch.full((B, T), 1.0 / (B * T), device=device)
        e = lambda *s: torch.empty(*s, device=device)  # noqa: E731
        # bf16 fused policy: the outputs of a (b,t) share one 1 KB row (include/rssm_rollout.h, RssmMtrssmOutputs.ld_*), as
        # rollout_ops.mtrssm_rollout allocates them; the other policies keep one dense tensor per output
        self.grouped = self.fused if grouped is None else grouped
        saved = torch.empty(_lib.mtrssm_saved_rows(B, precision), T, _lib.mtrssm_saved_elems(precision), device=device,
                            dtype=_lib.record_dtype(precision))
        if self.grouped:
            self.row, self.kl = e(B, T, _lib.MT_ROW_PITCH), e(B, T, 2)
            off = _lib.MT_ROW_OFFSETS
            width = lambda k: 96 if k == "feature" else 32 if k.startswith("hidden") else 16  # noqa: E731
            self.out = {k: self.row[..., o:o + width(k)] for k, o in off.items() if prior_sample or not k.startswith("prior_stoch")}
            self.out.update(kl_l=self.kl[..., 0], kl_h=self.kl[..., 1], saved=saved)
        else:
            self.out = {
                "feature": e(B, T, 96), "hidden_h": e(B, T, 32), "hidden_l": e(B, T, 32),
                "prior_probs_h": e(B, T, 8, 2), "prior_probs_l": e(B, T, 4, 4), "post_probs_h": e(B, T, 8, 2),
                "post_probs_l": e(B, T, 4, 4), "kl_l": e(B, T), "kl_h": e(B, T), "saved": saved,
            }
            if prior_sample:  # the prior MTState's own draws (mmtrssm/state.py:48-49): what the model API launches
                self.out["prior_stoch_h"], self.out["prior_stoch_l"] = e(B, T, 16), e(B, T, 16)
        self.gin = {
            "d_actions": e(B, T, 6), "d_embed_a": e(B, T, EW), "d_embed_v": e(B, T, EW), "d_deter_h0": e(B, 32), "d_deter_l0": e(B, 32),
            "d_hidden_h0": e(B, 32), "d_hidden_l0": e(B, 32), "d_stoch_h0": e(B, 16), "d_stoch_l0": e(B, 16),
            "dpre": torch.empty(B, T, _lib.MTRSSM_DPRE_FLOATS, device=device, dtype=_lib.record_dtype(precision)),
        }
        sizes = [w.numel() for w in self.weights]
        # two gradient buckets: with N > 1 the NCCL allreduce of step i's bucket runs on NCCL's stream under step i+1's kernels
        self.flat_grads = [torch.zeros(sum(sizes), device=device) for _ in range(2)]
        self.flat_grad = self.flat_grads[0]
        gws2 = [[g_.view_as(w) for g_, w in zip(f.split(sizes), self.weights)] for f in self.flat_grads]
        self.gws = gws2[0]
        L = _lib
        self.c_dims = L.MtrssmDims(B=B, T=T, A=6, E=64, HD=32, LD=32, HH=32, HR=32, CL=4, KL=4, CH=8, KH=2, l_tau=2.0, h_tau=4.0,
                                   precision=precision, obs_projected=int(obs_projected))
        fill = lambda st, d: [setattr(st, k, L.ptr(v)) for k, v in d.items()] and st  # noqa: E731
        self.c_w = fill(L.MtrssmWeights(), dict(zip(L.MT_WEIGHT_FIELDS, self.weights)))
        self.c_gws = [fill(L.MtrssmWeightGrads(), dict(zip(L.MT_WEIGHT_FIELDS, g_))) for g_ in gws2]
        self.c_gw = self.c_gws[0]
        self.c_in = fill(L.MtrssmInputs(), self.inp)
        self.c_out = L.MtrssmOutputs()
        for k, v in self.out.items():
            setattr(self.c_out, k, v.data_ptr())
        if self.grouped:
            self.c_out.ld_feature = self.c_out.ld_hidden = self.c_out.ld_probs = self.c_out.ld_stoch = L.MT_ROW_PITCH
            self.c_out.ld_kl = 2
        self.c_up = fill(L.MtrssmUpstream(), {"d_feature": self.d_feature, "d_kl_l": self.d_kl, "d_kl_h": self.d_kl})
        self.c_up.kl_wq, self.c_up.kl_wp = 0.2, 0.8
        self.c_gin = fill(L.MtrssmInputGrads(), self.gin)

    def fwd(self) -> None:
        self.lib.call("rssm_mtrssm_rollout_fwd", self.c_dims, self.c_w, self.c_in, self.c_out)

    def bwd_data(self) -> None:
        self.lib.call("rssm_mtrssm_rollout_bwd", self.c_dims, self.c_w, self.c_in, self.c_out, self.c_up, self.c_gin, None)

    def bwd_fused(self, k: int = 0, zero: bool = True) -> None:
        """bf16 path: BPTT + weight gradients in one kernel (no dpre round trip); gradients into bucket k (zeroed here unless the
        caller already did)."""
        if zero:
            self.flat_grads[k].zero_()
        self.lib.call("rssm_mtrssm_rollout_bwd", self.c_dims, self.c_w, self.c_in, self.c_out, self.c_up, self.c_gin, self.c_gws[k])

    def wgrad(self, k: int = 0) -> None:
        self.flat_grads[k].zero_()
        self.lib.call("rssm_mtrssm_wgrad", self.c_dims, self.c_in, self.c_out, self.gin["dpre"].data_ptr(), self.c_gws[k])

    def input_bytes(self) -> int:
        return sum(v.numel() * 4 for v in self.inp.values())


def time_direct(run: DirectMtrssm, steps: int, warmup: int, world: int) -> dict:
    import torch.distributed as dist

    # N > 1: ONE exchange per step, the mean of the 66 KB weight-gradient bucket, on the compute stream right after the backward.
    # Preferred: our one-shot kernel over NVLink peer memory (dp.P2pGradAllreduce); if the GPUs cannot map each other, NCCL.
    # (An asynchronous NCCL allreduce under the next step was measured and is WORSE at 8 GPUs, profiles/r2_u_bench_8gpu.json:
    # the collective's CTAs displace CTAs of the persistent rollout kernels, which then finish late by the whole delay.)
    count = [0]
    comm, reduced, mode = None, None, "none"
    if world > 1:
        from multimodal_mtrssm_b200 import dp

        try:
            comm = dp.P2pGradAllreduce(run.flat_grads[0].numel())
            reduced = torch.empty_like(run.flat_grads[0])
            mode = "one-shot allreduce kernel over NVLink peer memory (rssm_p2p_allreduce_mean)"
        except Exception as e:  # noqa: BLE001  (no peer mapping on this box / container: every rank raises together, see dp.py)
            comm = None
            mode = f"NCCL allreduce on the compute stream (peer mapping unavailable: {str(e)[:120]})"

    def step(evs=None):
        k = count[0] & 1
        if run.fused:  # the step's gradient bucket is cleared at the top of the step: part of the step, not of a kernel's own interval
            run.flat_grads[k].zero_()
        if evs:
            evs[0].record()
        run.fwd()
        if evs:
            evs[1].record()
        if run.fused:
            run.bwd_fused(k, zero=False)
        else:
            run.bwd_data()
        if evs:
            evs[2].record()
        if not run.fused:
            run.wgrad(k)
        if comm is not None:
            comm.allreduce(count[0], run.flat_grads[k], reduced)
        elif world > 1:
            dist.all_reduce(run.flat_grads[k])
        count[0] += 1
        if evs:
            evs[3].record()

    def drain():
        if comm is not None:
            comm.check()  # synchronises; raises if a rank gave up waiting for a peer

    for _ in range(warmup):
        step()
    drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(steps):
        step(evs[i])
    drain()  # every allreduce of the timed steps completes inside the timed region
    end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = start.elapsed_time(end)
    if world > 1:
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t)
    seg = [[e[i].elapsed_time(e[i + 1]) for e in evs] for i in range(3)]
    if comm is not None:  # the peer result equals NCCL's on the last step's bucket (checked outside the timed region)
        ref = run.flat_grads[(count[0] - 1) & 1].clone()
        dist.all_reduce(ref)
        ref /= world
        err = float((ref - reduced).abs().max() / ref.abs().max().clamp_min(1e-30))
        if err > 1e-5:
            raise SystemExit(f"p2p allreduce disagrees with NCCL: relative error {err}")
        torch.cuda.synchronize()
        dist.barrier()
        comm.close()
    return {"total_ms": total_ms, "fwd_ms": statistics.mean(seg[0]), "bwd_ms": statistics.mean(seg[1]),
            "wgrad_ms": statistics.mean(seg[2]), "allreduce": mode}


def time_e2e(B: int, T: int, precision: int, steps: int, warmup: int, world: int, device: torch.device, host_bf16: bool = False,
             prior_sample: bool = True, obs_projected: bool = False) -> dict:
    """Same metric through the public API with HOST buffers, as a training step sees it: the step's encoder outputs,
    actions and initial state are copied from pinned host memory (H2D), the noise is drawn on the device (as
    MoPoE_MMTRSSM.rollout_representation does), the rollout + autograd run through `rollout_ops.mtrssm_rollout`, the
    upstream gradient comes from a device-resident readout (stand-in for the decoders), and the loss + flat weight
    gradient are read back (D2H) every step."""
    import torch.distributed as dist

    from multimodal_mtrssm_b200 import rollout_ops as R
    from multimodal_mtrssm_b200 import synthetic
    from multimodal_mtrssm_b200.params import mtrssm_weight_list

    params = {k: v.to(device).requires_grad_(True) for k, v in synthetic.mtrssm_params().items()}
    weights = mtrssm_weight_list(params)
    batch = synthetic.mtrssm_batch(B, T, prior_noise=prior_sample)
    if obs_projected:  # side measurement only: the host holds the pre-multiplied partials (what a merged encoder layer emits)
        cpu_params = synthetic.mtrssm_params()
        for m, key in (("audio", "embed_a"), ("vision", "embed_v")):
            batch[key] = (batch[key] @ cpu_params[f"{m}_representation.rnn_to_post_projector.0.weight"][:, 32:].T).contiguous()
    # host_bf16 (side measurement only): the two embedding tensors wait on the host in bf16 -- what an autocast encoder emits and
    # what the bf16 policy's contractions consume anyway (the op widens them on the device) -- halving the PCIe bytes
    half = lambda k, v: v.bfloat16() if host_bf16 and k.startswith("embed_") else v  # noqa: E731
    host = {k: half(k, v).pin_memory() for k, v in batch.items() if not k.startswith("u_")}
    noise_shapes = {k: tuple(v.shape) for k, v in batch.items() if k.startswith("u_")}
    g = torch.Generator().manual_seed(7)
    readout = torch.randn(96, generator=g).to(device)  # device-resident "decoder": loss = <feature, readout> + KL terms
    n_w = sum(w.numel() for w in weights)
    host_out = torch.empty(n_w + 1).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    from multimodal_mtrssm_b200 import dp

    pre = dp.PinnedPrefetcher(host, device)  # step i+1's H2D copy runs on a side stream under step i's kernels

    def step(more: bool) -> None:
        slot, dev = pre.next()
        if more:
            pre.submit(host)
        dev = dict(dev)
        dev.update({k: torch.rand(s, device=device) for k, s in noise_shapes.items()})
        out = R.mtrssm_rollout(weights, precision=precision, obs_projected=obs_projected, **dev)
        loss = (out["feature"] @ readout).sum() + out["kl_l"].mean() + out["kl_h"].mean()
        grads = torch.autograd.grad(loss, weights)
        pre.release(slot)
        flat = torch.cat([loss.detach().reshape(1), *[g_.reshape(-1) for g_ in grads]])
        if world > 1:
            dist.all_reduce(flat)
        host_out.copy_(flat, non_blocking=True)

    def run(k: int) -> None:  # k steps, k H2D copies of the inputs (the first one is exposed), k D2H read-backs
        pre.submit(host)
        for i in range(k):
            step(i + 1 < k)

    run(max(1, warmup // 2))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(3, steps // 2)
    start.record()
    run(n)
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return {"value": world * B * T * n / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": (n_w + 1) * 4,
            "ms_per_step": ms / n, "steps": n, "gbs_per_rank": h2d * n / (ms * 1e-3) / 1e9,
            "note": "inputs (actions, both embeddings, initial state) copied from pinned host memory EVERY step (double-buffered: "
                    "step i+1's copy overlaps step i's kernels); noise drawn on the device; PCIe-bound"}


def time_mrssm(B: int, T: int, precision: int, device: torch.device, iters: int = 10) -> dict:
    """MoPoE-MRSSM (GRU transition; default.yaml sizes: the cfg1 model) rollout fwd+bwd through the public op API
    (`rollout_ops.mrssm_rollout` + autograd: forward kernel, BPTT kernel, weight-gradient kernel, torch allocations)."""
    from multimodal_mtrssm_b200 import rollout_ops as R
    from multimodal_mtrssm_b200 import synthetic
    from multimodal_mtrssm_b200.params import mrssm_weight_list

    weights = mrssm_weight_list({k: v.to(device).requires_grad_(True) for k, v in synthetic.mrssm_params().items()})
    inp = {k: v.to(device) for k, v in synthetic.mrssm_batch(B, T).items()}
    g = torch.Generator().manual_seed(7)
    readout = torch.randn(48, generator=g).to(device)

    def step() -> None:
        out = R.mrssm_rollout(weights, precision=precision, **inp)
        loss = (out["feature"] @ readout).sum() + out["kl"].mean()
        torch.autograd.grad(loss, weights)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(iters):
        step()
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end) / iters
    return {"workload": f"cfg1 model (MoPoE-MRSSM default.yaml sizes) B={B}", "B": B, "T": T, "ms_per_step": ms,
            "value": B * T / (ms * 1e-3), "via": "public op API (incl. torch allocations and the loss)"}


def time_train_step(B: int, T: int, world: int, device: torch.device, autocast: bool, iters: int = 10, graphed: bool = False) -> dict:
    """BASELINE.json's second metric, train sequences/s: one full training step of MoPoE-MMTRSSM built from the reference's
    default.yaml through this package's drop-in classes -- CNN encoders (stand-ins for the absent `cnn` package), initial
    state, fused rollout, decoders, Gaussian likelihood + KL, backward, flat-bucket gradient allreduce (N > 1), clip, AdamW.
    B sequences per GPU of synthetic audio / vision frames [B,T,1,32,32] ~ U(-1,1)."""
    from multimodal_mtrssm_b200 import compat, dp, standins, synthetic

    # Lightning's Trainer (the reference's driver, benchmark=None, not deterministic) turns cuDNN autotuning on: same here
    # (the stand-in encoders / decoders are cuDNN convolutions: 18.1 -> 10.6 ms per B = 256 step)
    torch.backends.cudnn.benchmark = True
    model = compat.load_model(ROOT / "multimodal_mtrssm_b200" / "configs" / "mopoe_mmtrssm_default.yaml")
    standins.materialize(model, model.feature_dim)
    model.to(device).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, capturable=graphed)
    # eager step: readiness-ordered buckets (the decoders' and the rollout's allreduce run under the encoders' backward);
    # graph step: one flat bucket inside the replay
    bucket = dp.FlatGradBucket(model.parameters()) if graphed or world == 1 else dp.OverlappedGradBuckets(model.parameters())
    g = torch.Generator().manual_seed(1234)
    frames = lambda: (torch.rand(B, T, 1, 32, 32, generator=g) * 2 - 1).to(device)  # noqa: E731
    act = synthetic.actions(B, T, g).to(device)
    batch = (act, frames(), frames(), act.clone(), frames(), frames())

    if graphed:  # the whole step (enc, rollout, dec, likelihood, backward, allreduce, clip, AdamW) as ONE CUDA graph replay
        gstep = dp.GraphedTrainStep(model, batch, opt, bucket, autocast_dtype=torch.bfloat16 if autocast else None)
        fresh = tuple(t.clone() for t in batch)  # a new batch every step: the copy into the graph's inputs is part of the step

        def step() -> None:
            gstep(fresh)
    else:
        def step() -> None:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                dp.train_step(model, batch, opt, bucket)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(iters):
        step()
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end) / iters
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return {"B_per_gpu": B, "T": T, "autocast_bf16": autocast, "cuda_graph": graphed, "ms_per_step": ms,
            "train_seq_per_sec": world * B / (ms * 1e-3)}


# ---------------------------------------------------------------------------------------------------
def time_wide_cfg3(device: torch.device, iters: int = 10, B: int = 1024, T: int = 64, D: int = 512) -> dict:
    """BASELINE.json cfg3: MoPoE-MRSSM rollout-only microbench, batch 1024, T = 64, hidden 512 -- fwd+bwd through the public op API
    (`rollout_ops.mrssm_rollout` + autograd: weight packing, persistent tcgen05 forward kernel, pre-pass, persistent BPTT kernel,
    embedding- and weight-gradient kernels, torch allocations).  Tensor-pipe bound: SURVEY.md 8(d) counts 5 445 632 forward
    GEMM FLOPs per (b,t), x3 for fwd+bwd; the fraction is against the measured sustained bf16 peak."""
    from multimodal_mtrssm_b200 import _lib
    from multimodal_mtrssm_b200 import rollout_ops as R
    from multimodal_mtrssm_b200 import synthetic
    from multimodal_mtrssm_b200.params import mrssm_weight_list

    weights = mrssm_weight_list({k: v.to(device).requires_grad_(True) for k, v in synthetic.mrssm_params(D=D).items()})
    inp = {k: v.to(device) for k, v in synthetic.mrssm_batch(B, T, D=D).items()}
    g = torch.Generator().manual_seed(7)
    readout = torch.randn(D + 16, generator=g).to(device)

    def run(grad: bool) -> float:
        def step() -> None:
            if grad:
                out = R.mrssm_rollout(weights, precision=_lib.PRECISION_BF16, **inp)
                torch.autograd.grad((out["feature"] @ readout).sum() + out["kl"].mean(), weights)
            else:
                with torch.no_grad():
                    R.mrssm_rollout(weights, precision=_lib.PRECISION_BF16, **inp)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(iters):
            step()
        end.record()
        torch.cuda.synchronize()
        return start.elapsed_time(end) / iters

    ms, ms_fwd = run(True), run(False)
    flops_fwd = 2 * ((6 + 16) * D + D * D + 3 * D * D + 3 * D * D + D * D + D * 16 + 2 * ((D + 64) * D + D * 16)) * B * T
    peaks_file = ROOT / "MEASURED_PEAKS.json"  # sustained bf16 figure: the kernels run inside a long step
    peak = json.loads(peaks_file.read_text()).get("bf16_tflops_sustained", 1384.6) if peaks_file.exists() else 1384.6
    return {"workload": f"cfg3 MoPoE-MRSSM rollout-only B={B} T={T} hidden={D}", "B": B, "T": T, "ms_per_step": ms, "fwd_only_ms": ms_fwd,
            "value": B * T / (ms * 1e-3), "unit": UNIT, "dtype": "bf16 operands, fp32 accumulate/state",
            "roofline": {"bound": "tensor", "achieved": 3 * flops_fwd / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": 3 * flops_fwd / (ms * 1e-3) / 1e12 / peak,
                         "serial_chain_note": f"{T} steps x 8 dependent contraction phases; {ms / T * 1e3:.1f} us per step"},
            "via": "public op API (incl. packing kernels, torch allocations and the loss)"}


def time_likelihood(device: torch.device, peak_gbs: float, B: int = 4096, T: int = 30, iters: int = 20) -> dict:
    """the fused Gaussian reconstruction likelihood of both modalities (objective.py:7-23 as
    mopoe_mrssm/core.py:294-303 calls it) on decoder-shaped tensors [B,T,1,32,32]; one launch forward, one backward.
    Algorithmic bytes: forward reads prediction + target (8 B per element), backward reads both and writes d prediction (12 B).
    4 x 503 MB of inputs exceed the 126 MB L2.  Beside it: the reference's own formulation (torch.distributions) eager on this GPU."""
    import torch.distributions as td

    from multimodal_mtrssm_b200 import objective

    g = torch.Generator(device=device).manual_seed(1234)
    mk = lambda: torch.rand(B, T, 1, 32, 32, device=device, generator=g) * 2 - 1  # noqa: E731
    preds, tgts = [mk().requires_grad_(True), mk().requires_grad_(True)], [mk(), mk()]
    n = 2 * B * T * 1024
    ones = torch.ones(2, device=device)

    def timed(fn) -> float:  # noqa: ANN001
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(iters):
            fn()
        end.record()
        torch.cuda.synchronize()
        return start.elapsed_time(end) / iters

    det = [p.detach() for p in preds]
    nb = [B * T, B * T]
    fwd_ms = timed(lambda: objective.gaussian_nll_op(det, tgts, nb, 1.0))
    bwd_ms = timed(lambda: objective.gaussian_nll_bwd_op(det, tgts, nb, 1.0, ones, False))

    def ours() -> None:
        torch.autograd.grad(objective.likelihood_pairs(preds, tgts, 3).sum(), preds)

    def ref() -> None:
        loss = sum(-td.Independent(td.Normal(p, 1.0), 3).log_prob(t).mean() for p, t in zip(preds, tgts))
        torch.autograd.grad(loss, preds)

    ours_ms, ref_ms = timed(ours), timed(ref)
    return {"workload": f"both modalities [B={B},T={T},1,32,32] fp32, fwd+bwd", "elements": n,
            "fwd_kernel_ms": fwd_ms, "bwd_kernel_ms": bwd_ms, "fwd_bwd_via_api_ms": ours_ms,
            "roofline_fwd": {"bound": "hbm", "achieved": 8 * n / (fwd_ms * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                             "frac": 8 * n / (fwd_ms * 1e-3) / 1e9 / peak_gbs, "algorithmic_bytes_per_launch": 8 * n},
            "roofline_bwd": {"bound": "hbm", "achieved": 12 * n / (bwd_ms * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                             "frac": 12 * n / (bwd_ms * 1e-3) / 1e9 / peak_gbs, "algorithmic_bytes_per_launch": 12 * n},
            "reference_formulation_eager_on_this_gpu_ms": ref_ms, "speedup_vs_reference_eager_gpu": ref_ms / ours_ms}


def reference_eager_gpu_cfg3(device: torch.device, B: int = 1024, T: int = 64, D: int = 512) -> dict:
    """The "vs reference" number of cfg3: the fp32 oracle (= the reference's PyTorch rollout with explicit noise)
    run EAGER on this GPU, fwd + autograd bwd, CUDA events.  Baseline leg: touches oracle/."""
    from multimodal_mtrssm_b200 import synthetic
    from oracle import rssm_oracle as O

    params = {k: v.to(device).requires_grad_(True) for k, v in synthetic.mrssm_params(D=D).items()}
    inp = {k: v.to(device) for k, v in synthetic.mrssm_batch(B, T, D=D).items()}
    up = torch.randn(B, T, D + 16, generator=torch.Generator().manual_seed(7)).to(device)

    def step() -> None:
        res = O.mrssm_rollout(params, C=4, K=4, u_prior=None, **inp)
        loss = (res["post_feature"] * up).sum() + O.kl_per_sample(res["post_probs"], res["prior_probs"], True).mean()
        torch.autograd.grad(loss, list(params.values()))

    step()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(3):
        step()
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end) / 3
    return {"ms_per_step": ms, "value": B * T / (ms * 1e-3), "unit": UNIT, "what": "fp32 PyTorch oracle, eager on this GPU (cuBLAS fp32 / TF32 off)"}


def reference_eager_gpu_cfg2(device: torch.device, B: int, T: int) -> dict:
    """The same comparison for the headline workload (cfg2 sizes): the fp32 oracle (= the reference's MoPoE-MMTRSSM rollout with
    explicit noise: the Python T loop of ~100 small launches per step and direction) EAGER on this GPU, fwd + autograd bwd, on the
    SAME batch, CUDA events.  This is what a reference user gets from a B200 today.  Baseline leg: touches oracle/."""
    from multimodal_mtrssm_b200 import synthetic
    from oracle import rssm_oracle as O

    dims = dict(CL=4, KL=4, CH=8, KH=2, l_tau=2.0, h_tau=4.0)
    params = {k: v.to(device).requires_grad_(True) for k, v in synthetic.mtrssm_params().items()}
    inp = {k: v.to(device) for k, v in synthetic.mtrssm_batch(B, T).items()}
    up = torch.randn(B, T, 96, generator=torch.Generator().manual_seed(7)).to(device)

    def step() -> None:
        res = O.mtrssm_rollout(params, dims=dims, u_prior_l=None, u_prior_h=None, **inp)
        kl_l = O.kl_per_sample(res["post_probs_l"], res["prior_probs_l"], True).mean()
        kl_h = O.kl_per_sample(res["post_probs_h"], res["prior_probs_h"], True).mean()
        loss = (res["post_feature"] * up).sum() + kl_l + kl_h
        torch.autograd.grad(loss, list(params.values()), allow_unused=True)

    step()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(3):
        step()
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end) / 3
    return {"B": B, "T": T, "ms_per_step": ms, "value": B * T / (ms * 1e-3), "unit": UNIT,
            "what": "fp32 PyTorch oracle of the same workload, eager on this GPU"}


# CPU legs (the ONLY place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------------------
def cpu_oracle_rate(B: int, T: int, literal: bool, budget_s: float) -> dict:
    """Times the CPU oracle (fp32, all host threads) on a bounded sample of the same workload: fwd + autograd bwd."""
    from multimodal_mtrssm_b200 import synthetic
    from oracle import rssm_oracle as O

    dims = dict(CL=4, KL=4, CH=8, KH=2, l_tau=2.0, h_tau=4.0)
    params = {k: v.clone().requires_grad_(True) for k, v in synthetic.mtrssm_params().items()}
    inp = synthetic.mtrssm_batch(B, T, prior_noise=True)
    g = torch.Generator().manual_seed(7)
    up = torch.randn(B, T, 96, generator=g)

    def step() -> None:
        if literal:
            x = {k: v for k, v in inp.items() if not k.startswith("u_")}
            res = O.mtrssm_rollout_literal(params, dims=dims, **x)
        else:
            res = O.mtrssm_rollout(params, dims=dims, **inp)
        kl_l = O.kl_per_sample(res["post_probs_l"], res["prior_probs_l"], True).mean()
        kl_h = O.kl_per_sample(res["post_probs_h"], res["prior_probs_h"], True).mean()
        loss = (res["post_feature"] * up).sum() + kl_l + kl_h
        torch.autograd.grad(loss, list(params.values()), allow_unused=True)

    step()  # warm-up
    t0 = time.perf_counter()
    step()
    one = time.perf_counter() - t0
    reps = max(1, min(10, int(budget_s / max(one, 1e-3))))
    times = [one]
    for _ in range(reps - 1):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    best = statistics.median(times)
    return {"value": B * T / best, "unit": UNIT, "cores": torch.get_num_threads(), "ms_per_step": best * 1e3,
            "sample": f"oracle {'literal (reference structure, global RNG)' if literal else 'functional'} fp32 fwd+autograd bwd, "
                      f"B={B} T={T} of the same workload, median of {len(times)}"}


def workload_config(B: int, T: int, world: int, prior_sample: bool) -> dict:
    """`config` of the JSON line: identical in both arms (the reference arm times the same workload on a bounded sample)."""
    input_mb = ((6 + 2 * 64 + 12 + (12 if prior_sample else 0)) * T + 160) * 4 * B / 1e6  # actions, embeddings, uniforms, initial state
    cfg = {
        "workload": "cfg2: MoPoE-MMTRSSM default.yaml sizes (hd=ld=32, hs=ls=16, E=64, A=6), rollout fwd+bwd on synthetic "
                    "vision+audio embeddings/actions", "batch_per_gpu": B, "seq_len": T, "global_batch": world * B,
        "parallelism": f"dp{world} (batch-sharded, one mean-allreduce of the flat weight-gradient bucket per step)" if world > 1 else "single GPU",
        "prior_samples": "drawn and written (what MoPoE_MMTRSSM.rollout_representation launches)" if prior_sample else "not drawn (bytes subtracted)",
    }
    cfg["l2"] = f"inputs {input_mb:.0f} MB + outputs/records per step exceed the 126 MB L2 (no flush needed)"
    return cfg


def run_reference(args: argparse.Namespace) -> None:
    """Reference arm: the reference's own CPU formulation of the path (the oracle's LITERAL variant: Python T loop, State-style
    per-step draws from the global RNG, per-step cats and T-way stacks) on ALL host cores.  `/root/reference` itself cannot be
    imported or installed (lightning / torchrl / distribution_extension / cnn absent, hatchling missing: DESIGN.md), so `kind` is
    "port".  Every step is a bounded sample of the SAME workload (same dims, same T, batch B_s <= B sized from a probe so that
    W + K steps fit `--ref-budget-s`); the rate is also reported at B = 256 so the batch dependence is visible."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(ncpu)  # torchrun exports OMP_NUM_THREADS=1: undo it, this arm uses every host core
    T, K, W = args.seq_len, max(1, args.steps), max(1, args.warmup)
    probe = cpu_oracle_rate(256, T, literal=True, budget_s=2.0)
    per_step_s = args.ref_budget_s * 0.85 / (K + W)
    Bs = int(probe["value"] * per_step_s / T) // 256 * 256
    Bs = max(256, min(args.batch, Bs))
    from multimodal_mtrssm_b200 import synthetic
    from oracle import rssm_oracle as O

    dims = dict(CL=4, KL=4, CH=8, KH=2, l_tau=2.0, h_tau=4.0)
    params = {k: v.clone().requires_grad_(True) for k, v in synthetic.mtrssm_params().items()}
    inp = {k: v for k, v in synthetic.mtrssm_batch(Bs, T).items() if not k.startswith("u_")}
    up = torch.randn(Bs, T, 96, generator=torch.Generator().manual_seed(7))

    def step() -> None:
        res = O.mtrssm_rollout_literal(params, dims=dims, **inp)
        kl_l = O.kl_per_sample(res["post_probs_l"], res["prior_probs_l"], True).mean()
        kl_h = O.kl_per_sample(res["post_probs_h"], res["prior_probs_h"], True).mean()
        torch.autograd.grad((res["post_feature"] * up).sum() + kl_l + kl_h, list(params.values()), allow_unused=True)

    t_start = time.perf_counter()
    for _ in range(W):
        step()
    times = []
    for _ in range(K):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > args.ref_budget_s * 1.5:  # a slower box than the probe suggested: stop early, say so
            break
    total = sum(times)
    value = Bs * T * len(times) / total
    sample = (f"oracle literal (reference structure, global RNG) fp32 fwd+autograd bwd on {ncpu} host threads; every step = B_s={Bs} "
              f"sequences of the B={args.batch} workload, {W} warm-up + {len(times)} timed steps; at B=256: {probe['value']:.3e} steps/s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": W, "ms_per_step": total / len(times) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch, T, args.gpus, not args.no_prior_sample),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncpu, "kind": "port", "sample": sample,
                         "rate_at_B256": probe["value"], "sample_batch": Bs},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference cannot be imported here (lightning/torchrl/distribution_extension absent, SURVEY.md §8(c)); "
                "this is the oracle restatement with the reference's per-step structure on the host cores",
    }
    print(json.dumps(line))


def measured_traffic(stamp: str, kernel: str, B: int, T: int, prior_sample: bool) -> float | None:
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture of THIS build
    (profiles/traffic.json: {"build_stamp", "batch", "seq_len", "prior_sample", "kernels": {name: bytes}}); None when the library
    was rebuilt since (stamp mismatch) or the workload differs -- never a stale constant."""
    f = ROOT / "profiles" / "traffic.json"
    if not f.exists():
        return None
    try:
        d = json.loads(f.read_text())
    except ValueError:
        return None
    if d.get("build_stamp") != stamp or (d.get("batch"), d.get("seq_len"), d.get("prior_sample")) != (B, T, prior_sample):
        return None
    return d.get("kernels", {}).get(kernel)


def main() -> None:
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    from multimodal_mtrssm_b200 import dp as _dp

    numa_bound = os.environ.get("RSSM_NO_NUMA_BIND") is None and _dp.bind_to_gpu_numa_node(local)  # before any pinned allocation: the e2e leg copies 633 MB per step and rank
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)
    from multimodal_mtrssm_b200 import _lib
    from multimodal_mtrssm_b200.build import build_stamp

    precision = {"bf16": _lib.PRECISION_BF16, "bf16_fused": _lib.PRECISION_BF16_FUSED, "fp32": _lib.PRECISION_FP32}[args.precision]
    B, T = args.batch, args.seq_len
    prior_sample = not args.no_prior_sample
    fwd_bytes_bt, step_bytes_bt = bytes_per_bt(prior_sample)
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peaks = json.loads(peaks_file.read_text())
        peak, peak_src = peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peaks, peak, peak_src = {}, 6650.0, "B200_PROFILING.md fallback (of fallback)"
    launches0 = _lib.launch_count()
    run = DirectMtrssm(B, T, precision, device, prior_sample)
    # nvidia-smi delivers its first sample after ~100 ms and a short timed region (K steps of ~2 ms) may end before that:
    # sample from before the warm-up until the end-to-end measurement is done (the GPU is under this load throughout)
    sampler = ClockSampler(local) if rank == 0 else None
    res = time_direct(run, args.steps, args.warmup, world)
    launches = (_lib.launch_count() - launches0) // (args.steps + args.warmup)
    e2e = time_e2e(B, T, precision, args.steps, args.warmup, world, device, prior_sample=prior_sample)
    if sampler and len(sampler.samples) < 3:  # still too short: keep the same kernels running for ~0.5 s
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end:
            run.fwd()
            torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None

    def frac_of_hbm(b_: int, t_: int, ms: float) -> float:
        return step_bytes_bt * b_ * t_ / (ms * 1e-3) / 1e9 / peak

    extras, summary = {}, {}
    if not args.no_extras:  # every rank takes part (gradient allreduce inside)
        train = [time_train_step(b_, T, world, device, ac, graphed=gr)
                 for b_, ac, gr in ((8, True, False), (8, True, True), (256, True, False), (256, True, True), (256, False, False))]
        if rank == 0:
            extras["train_step"] = {"metric": "train sequences/s (full MoPoE-MMTRSSM training step, default.yaml model)",
                                    "runs": train}
            best = max(train, key=lambda r_: r_["train_seq_per_sec"])
            summary["train_seq_per_sec"] = {"value": round(best["train_seq_per_sec"], 1), "n_gpus": world, "B_per_gpu": best["B_per_gpu"],
                                            "ms_per_step": round(best["ms_per_step"], 3), "cuda_graph": best["cuda_graph"]}
    if rank == 0 and not args.no_extras:
        n2 = max(3, args.steps // 2)
        for name, prec in (("fp32_path", _lib.PRECISION_FP32), ("bf16_two_kernel_backward_path", _lib.PRECISION_BF16),
                           ("bf16_fused_backward_path", _lib.PRECISION_BF16_FUSED)):
            if prec == precision:
                continue
            other = DirectMtrssm(B, T, prec, device, prior_sample)
            r2 = time_direct(other, n2, 3, 1)
            extras[name] = {"value": B * T * n2 / (r2["total_ms"] * 1e-3), "ms_per_step": r2["total_ms"] / n2,
                            "fwd_ms": r2["fwd_ms"], "bwd_ms": r2["bwd_ms"], "wgrad_ms": r2["wgrad_ms"],
                            "frac_of_hbm": frac_of_hbm(B, T, r2["total_ms"] / n2)}
            del other
        if "fp32_path" in extras:
            summary["fp32_path"] = {"ms_per_step": round(extras["fp32_path"]["ms_per_step"], 3), "frac_of_hbm": round(extras["fp32_path"]["frac_of_hbm"], 3),
                                    "parity": "1e-5 vs oracle (tests)"}
        small = DirectMtrssm(8, T, precision, device, prior_sample)
        r3 = time_direct(small, 50, 10, 1)
        extras["default_batch8"] = {"B": 8, "T": T, "value": 8 * T * 50 / (r3["total_ms"] * 1e-3), "us_per_step": r3["total_ms"] / 50 * 1e3}
        summary["default_yaml_B8_fwd_bwd_us"] = round(extras["default_batch8"]["us_per_step"], 1)
        del small
        # the other batch sizes SURVEY.md 8(d) asks for, and cfg4 (long horizon); same dims, same kernels
        extras["other_workloads"] = []
        for name, b_, t_ in (("cfg2 B=256", 256, T), ("cfg2 B=4096", 4096, T), ("cfg2 B=16384", 16384, T), ("cfg4 B=256 T=512", 256, 512),
                             ("one tile B=16 T=512 (serial chain)", 16, 512)):
            w = DirectMtrssm(b_, t_, precision, device, prior_sample)
            rr = time_direct(w, 10, 3, 1)
            extras["other_workloads"].append({"workload": name, "B": b_, "T": t_, "value": b_ * t_ * 10 / (rr["total_ms"] * 1e-3),
                                              "ms_per_step": rr["total_ms"] / 10, "fwd_ms": rr["fwd_ms"], "bwd_ms": rr["bwd_ms"] + rr["wgrad_ms"],
                                              "frac_of_hbm": frac_of_hbm(b_, t_, rr["total_ms"] / 10)})
            del w
        cfg4, chain = extras["other_workloads"][3], extras["other_workloads"][4]
        # serial-chain lower bound T x t_step_min: t_step_min = one 16-sequence tile alone on the GPU, fwd + bwd
        summary["cfg4_T512_B256"] = {"ms": round(cfg4["ms_per_step"], 3), "steps_per_s": round(cfg4["value"]), "frac_of_hbm": round(cfg4["frac_of_hbm"], 4),
                                     "serial_chain_bound_ms": round(chain["ms_per_step"], 3),
                                     "t_step_min_us": round(chain["ms_per_step"] / 512 * 1e3, 2)}
        mr_prec = _lib.PRECISION_FP32 if precision == _lib.PRECISION_FP32 else _lib.PRECISION_BF16
        extras["mrssm_workloads"] = [time_mrssm(8, T, mr_prec, device), time_mrssm(16384, T, mr_prec, device)]
        summary["cfg1_mrssm_B8_fwd_bwd_us"] = round(extras["mrssm_workloads"][0]["ms_per_step"] * 1e3, 1)
        if precision != _lib.PRECISION_FP32:  # the wide family is the bf16 tensor-core path
            extras["cfg3_wide_mrssm"] = time_wide_cfg3(device)
            extras["cfg3_wide_mrssm"]["reference_eager_on_this_gpu"] = reference_eager_gpu_cfg3(device)
            r = extras["cfg3_wide_mrssm"]
            r["speedup_vs_reference_eager_gpu"] = r["reference_eager_on_this_gpu"]["ms_per_step"] / r["ms_per_step"]
            summary["cfg3_B1024_T64_D512"] = {"ms": round(r["ms_per_step"], 3), "tflops": round(r["roofline"]["achieved"], 1),
                                              "frac_of_tensor": round(r["roofline"]["frac"], 3), "x_vs_eager_gpu": round(r["speedup_vs_reference_eager_gpu"], 1)}
        extras["reference_eager_on_this_gpu"] = reference_eager_gpu_cfg2(device, B, T)
        extras["reference_eager_on_this_gpu"]["speedup_of_value"] = (B * T / (res["total_ms"] / args.steps * 1e-3)) / extras["reference_eager_on_this_gpu"]["value"]
        summary["x_vs_oracle_eager_same_gpu_same_batch"] = round(extras["reference_eager_on_this_gpu"]["speedup_of_value"], 1)
        extras["e2e_bf16_host_embeddings"] = time_e2e(B, T, precision, 10, 4, 1, device, host_bf16=True, prior_sample=prior_sample)
        if precision == _lib.PRECISION_BF16_FUSED:  # SURVEY §8 f2 (MoPoE_MMTRSSM.hoist_obs_projection): pre-multiplied first-layer partials
            pj = DirectMtrssm(B, T, precision, device, prior_sample, obs_projected=True)
            rp = time_direct(pj, n2, 3, 1)
            del pj
            ep = time_e2e(B, T, precision, 10, 4, 1, device, prior_sample=prior_sample, obs_projected=True)
            extras["obs_projected_f2"] = {"ms_per_step": rp["total_ms"] / n2, "fwd_ms": rp["fwd_ms"], "bwd_ms": rp["bwd_ms"] + rp["wgrad_ms"],
                                          "value": B * T * n2 / (rp["total_ms"] * 1e-3), "e2e": ep,
                                          "note": "kernel boundary carries e @ W1[:, 32:].T (32 floats per modality instead of 64): the GEMM that "
                                                  "produces it (or the merged encoder layer) is outside this number"}
            summary["obs_projected_f2"] = {"ms_per_step": round(rp["total_ms"] / n2, 4), "e2e_steps_per_s": round(ep["value"])}
        extras["likelihood_f3"] = time_likelihood(device, peak)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    if rank != 0:
        return

    ms_step = res["total_ms"] / args.steps
    value = world * B * T / (ms_step * 1e-3)
    seg = {"mtrssm_fwd_kernel": res["fwd_ms"], "mtrssm_bwd_kernel": res["bwd_ms"], "wgrad_kernel": res["wgrad_ms"]}
    if run.fused:
        seg = {"mtrssm_fwd2_kernel": res["fwd_ms"], "mtrssm_bwd_fused2_kernel": res["bwd_ms"] + res["wgrad_ms"]}
    dominant = max(seg, key=seg.get)
    # algorithmic bytes of the dominant launch: forward = fwd I/O; backward (bwd kernel + its wgrad pass) = fwd I/O again
    if dominant.startswith("mtrssm_fwd"):
        alg_bytes, dur_ms, what = fwd_bytes_bt * B * T, seg[dominant], "forward rollout kernel"
    else:
        alg_bytes, dur_ms = fwd_bytes_bt * B * T, res["bwd_ms"] + res["wgrad_ms"]
        what = ("fused backward kernel (BPTT + weight gradients)" if run.fused
                else "backward = BPTT kernel + weight-gradient kernel (dominant: %s)" % dominant)
    achieved = alg_bytes / (dur_ms * 1e-3) / 1e9
    stamp = build_stamp()
    traffic = measured_traffic(stamp, dominant, B, T, prior_sample) if run.fused else None
    summary["kernel_ms"] = {k: round(v, 4) for k, v in seg.items()}
    summary["step_frac_of_hbm"] = round(step_bytes_bt * B * T / (ms_step * 1e-3) / 1e9 / peak, 4)
    summary["fwd_frac_of_hbm"] = round(fwd_bytes_bt * B * T / (res["fwd_ms"] * 1e-3) / 1e9 / peak, 4)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if precision == _lib.PRECISION_FP32 else "bf16", "data": "synthetic",
        "config": workload_config(B, T, world, prior_sample),
        "allreduce": res.get("allreduce"),
        "kernel_ms": seg,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": what, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                     "algorithmic_bytes_per_bt": fwd_bytes_bt, "build_stamp": stamp[:16]},
        "roofline_step": {"algorithmic_bytes": step_bytes_bt * B * T, "achieved_gbs": step_bytes_bt * B * T / (ms_step * 1e-3) / 1e9,
                          "frac_of_hbm": step_bytes_bt * B * T / (ms_step * 1e-3) / 1e9 / peak,
                          "tensor_tflops": 3 * FWD_FLOPS_PER_BT * B * T / (ms_step * 1e-3) / 1e12},
        "precision": "bf16 operands / fp32 accumulate+state (tensor-core path)" if precision != _lib.PRECISION_FP32 else "3-way bf16 split, fp32-level accuracy",
        "e2e": e2e | {"host_numa_bound": numa_bound}, "gpu_launches": int(launches), "clocks": clocks, **extras,
    }
    if not args.no_cpu_baseline and world == 1:
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        torch.set_num_threads(ncpu)
        cb = cpu_oracle_rate(args.cpu_batch, T, literal=False, budget_s=12.0)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "sample")} | {"kind": "port"}
    # the compact summary goes LAST (the driver keeps the tail of stdout) and, abridged, into `roofline` (the driver keeps that dict
    # whole; `config` stays identical to the reference arm's)
    line["summary"] = summary
    line["roofline"]["summary"] = {k: summary[k] for k in ("train_seq_per_sec", "cfg4_T512_B256", "cfg3_B1024_T64_D512", "fp32_path", "step_frac_of_hbm")
                                   if k in summary}
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
