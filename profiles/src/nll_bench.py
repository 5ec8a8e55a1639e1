"""Times the likelihood kernels alone (profiles/src/nll_bench.py [B] [T]); used for the ncu capture of profiles/r1_j_*."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 30
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
print(json.dumps(bench.time_likelihood(torch.device("cuda", 0), 6464.9, B, T, iters)))
