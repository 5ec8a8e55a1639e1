"""B200-native drop-in for the latent-rollout hot path of Mamo1031/Multimodal-MTRSSM.

Layers (bottom up):
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/rssm_rollout.h)
  _lib.py          ctypes binding of librssm_rollout.so (built in-tree by build.py; no fallback)
  rollout_ops.py   torch custom ops with autograd over the C ABI
  distribution.py, state.py, mtstate.py, networks.py, core.py, mopoe_mrssm.py, mopoe_mmtrssm.py, objective.py
                   host-side mirror of the reference's Python interface for this path
  compat.py        registers the mirror under the reference's module paths (`multimodal_rssm.models...`)
  dp.py            batch-sharded data parallelism: one flat-bucket gradient allreduce
"""

__version__ = "0.1.0"
