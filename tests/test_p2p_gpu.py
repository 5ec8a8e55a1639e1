"""One-shot peer-memory allreduce of the gradient bucket (csrc/p2p_allreduce.cu, dp.P2pGradAllreduce) against NCCL,
two ranks on two GPUs of one box.  Skipped on a single-GPU box (the driver's GPU test tier); run with `gpurun --gpus 2`."""

from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO"])
from multimodal_mtrssm_b200 import dp
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%s" % os.environ["PORT"], rank=rank, world_size=world,
                        device_id=torch.device("cuda", rank))
for n in (16512, 7, 4099):                       # the rollout's bucket, a sub-vector size, a ragged size
    comm = dp.P2pGradAllreduce(n)
    out = torch.empty(n, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1000 * rank + n)
    for step in range(6):                          # both slots, several epochs each; rank 1 lags on purpose at step 3
        b = comm.bucket(step)
        b.copy_(torch.randn(n, generator=g, device="cuda") * (step + 1))
        if rank == 1 and step == 3:
            torch.cuda._sleep(200_000_000)         # ~0.1 s of device time: rank 0 must wait at the flags, not read early
        if step % 3 == 2:                          # in place on ordinary memory (what the training step does)
            src = b.clone()
            comm.allreduce(step, src)
            out.copy_(src)
        else:                                      # bucket filled in place
            comm.allreduce(step, None, out)
        ref = b.clone()
        dist.all_reduce(ref)
        ref /= world
        torch.testing.assert_close(out, ref, rtol=1e-6, atol=1e-6)
    comm.check()
    dist.barrier()
    comm.close()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one box")
def test_p2p_allreduce_matches_nccl_world_size_2(tmp_path):
    script = tmp_path / "p2p_worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [
        subprocess.Popen([sys.executable, str(script)], env={**os.environ, "RANK": str(r), "WORLD_SIZE": "2", "PORT": port, "REPO": str(ROOT)},
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        for r in range(2)
    ]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
