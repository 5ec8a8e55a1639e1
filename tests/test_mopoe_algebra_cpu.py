"""The probability-domain form of the MoPoE fusion that the bf16 kernels evaluate (multimodal_mtrssm_b200/csrc/frag.cuh:
`mopoe_posterior_fast`, `mopoe_responsibilities_fast`) against the oracle's log-domain restatement of the
reference (`oracle.rssm_oracle.mopoe_fuse`, mrssm/mopoe_mrssm/core.py:241-251,135-154), in float64 on the CPU:

    pa = softmax_16(la), pv = softmax_16(lv), s = pa + pv + pa pv
    posterior  q  = softmax_group(mopoe_fuse(la, lv)) = s / sum_group(s)
    d mixed / d log_softmax(la) = (pa + pa pv) / s,   d mixed / d log_softmax(lv) = (pv + pa pv) / s      (the responsibilities)
"""

import pytest
import torch

from oracle import rssm_oracle as O


@pytest.mark.parametrize("K", [2, 4, 8, 16])
@pytest.mark.parametrize("scale", [1.0, 8.0, 30.0])
def test_probability_domain_posterior_equals_the_reference_fusion(K, scale):
    g = torch.Generator().manual_seed(K * 100 + int(scale))
    la = (torch.randn(257, 16, generator=g, dtype=torch.float64) * scale).requires_grad_(True)
    lv = (torch.randn(257, 16, generator=g, dtype=torch.float64) * scale).requires_grad_(True)
    mixed = O.mopoe_fuse(la, lv)
    q_ref = torch.softmax(mixed.reshape(-1, 16 // K, K), dim=-1).reshape(-1, 16)
    pa, pv = torch.softmax(la, -1), torch.softmax(lv, -1)
    s = pa + pv + pa * pv
    q = (s.reshape(-1, 16 // K, K) / s.reshape(-1, 16 // K, K).sum(-1, keepdim=True)).reshape(-1, 16)
    assert torch.allclose(q, q_ref, rtol=1e-10, atol=1e-300)
    # responsibilities = the Jacobian of `mixed` w.r.t. the two flat log-softmaxes (diagonal), checked through autograd: with an
    # upstream gradient dm on `mixed`, d la = ra dm - pa sum(ra dm) (flat log-softmax backward), likewise for lv
    dm = torch.randn(257, 16, generator=g, dtype=torch.float64)
    gla, glv = torch.autograd.grad((mixed * dm).sum(), (la, lv))
    ra, rv = (pa + pa * pv) / s, (pv + pa * pv) / s
    dla = ra * dm - pa * (ra * dm).sum(-1, keepdim=True)
    dlv = rv * dm - pv * (rv * dm).sum(-1, keepdim=True)
    assert torch.allclose(dla, gla, rtol=1e-9, atol=1e-12) and torch.allclose(dlv, glv, rtol=1e-9, atol=1e-12)
