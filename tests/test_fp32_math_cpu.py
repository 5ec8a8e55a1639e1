"""CPU restatement (numpy, float32) of the short fp32-accurate math the fp32-parity kernels use instead of libm
(multimodal_mtrssm_b200/csrc/frag.cuh: `Math<false>`, `split_pack<3>`), checked against float64:

* exp(x) = ex2(t) * (1 + e ln 2) with t + e = x log2(e) formed by a product, an FMA residual and the tail of the constant;
* tanh(x) = (1 - e) / (1 + e), e = exp(-2|x|), with the odd series below |x| = 0.1;
* ELU's exp(x) - 1 by its Taylor series above -0.1;
* the EXACT 3-way bf16 split by truncation: x = hi + mid + lo bit for bit (for |x| >= 2^-100: below that the residuals become
  subnormal and lose bits -- far outside the range of activations, weights and gradients).

The device MUFU units add at most 2^-22 relative (ex2 / lg2) and 1 ulp (rcp) to the errors of the ideal arithmetic modelled
here; the end-to-end statement is the 1e-5 parity of the GPU tests (tests/test_rollout_gpu.py, tests/test_bench_configs_gpu.py)."""

import numpy as np

f32 = np.float32
L2E_HI, L2E_LO, LN2 = f32(1.4426950216293335), f32(1.9259629911266175e-8), f32(0.6931471805599453)


def _ex2(t: np.ndarray) -> np.ndarray:  # an ideal ex2 unit: exact, rounded to float32
    return np.exp2(t.astype(np.float64)).astype(f32)


def _fma(a: np.ndarray, b, c: np.ndarray) -> np.ndarray:  # one rounding
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(f32)


def exp_short(x: np.ndarray) -> np.ndarray:
    x = x.astype(f32)
    t = (x * L2E_HI).astype(f32)
    e = (_fma(x, L2E_HI, -t) + (x * L2E_LO).astype(f32)).astype(f32)
    y = _ex2(t)
    return _fma(y, 1.0, (y * (e * LN2).astype(f32)).astype(f32))


def tanh_short(x: np.ndarray) -> np.ndarray:
    x = x.astype(f32)
    ax, x2 = np.abs(x), (x * x).astype(f32)
    e = exp_short(-2 * ax)
    big = ((1 - e).astype(f32) / (1 + e).astype(f32)).astype(f32)
    small = (ax * (1 + x2 * (f32(-0.33333334) + x2 * f32(0.13333333)))).astype(f32)
    return np.copysign(np.where(ax < 0.1, small, big), x).astype(f32)


def expm1_short(x: np.ndarray) -> np.ndarray:
    x = x.astype(f32)
    series = (x * (1 + x * (0.5 + x * (f32(0.16666667) + x * (f32(4.1666668e-2) + x * f32(8.3333333e-3)))))).astype(f32)
    return np.where(x > -0.1, series, (exp_short(x) - 1).astype(f32)).astype(f32)


def test_exp_is_accurate_to_two_ulp_over_the_range_the_kernels_see():
    x = np.linspace(-40, 10, 400001).astype(f32)
    ref = np.exp(x.astype(np.float64))
    rel = np.abs(exp_short(x).astype(np.float64) - ref) / ref
    assert rel.max() < 2.4e-7, rel.max()  # 2 ulp of float32 = 2.4e-7


def test_tanh_is_accurate_in_relative_terms_down_to_zero():
    x = np.concatenate([np.linspace(-8, 8, 400001), np.logspace(-8, -1, 2001), -np.logspace(-8, -1, 2001)]).astype(f32)
    ref = np.tanh(x.astype(np.float64))
    err = np.abs(tanh_short(x).astype(np.float64) - ref)
    assert err.max() < 1.5e-7, err.max()
    nz = np.abs(ref) > 0
    assert (err[nz] / np.abs(ref[nz])).max() < 4e-7


def test_expm1_of_elu_does_not_cancel_near_zero():
    x = np.concatenate([np.linspace(-12, 0, 200001), -np.logspace(-8, -1, 2001)]).astype(f32)
    ref = np.expm1(x.astype(np.float64))
    err = np.abs(expm1_short(x).astype(np.float64) - ref)
    assert err.max() < 1.2e-7, err.max()
    nz = np.abs(ref) > 1e-30
    assert (err[nz] / np.abs(ref[nz])).max() < 5e-7


def test_truncation_split_is_exact_and_every_term_is_a_bf16():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(200000) * 10.0 ** rng.integers(-6, 6, 200000), [0.0, 1.0, -1.0, 3.4e38, 1.0e-30]]).astype(f32)
    mask = np.uint32(0xFFFF0000)
    hi = (x.view(np.uint32) & mask).view(f32)
    r = (x - hi).astype(f32)
    mid = (r.view(np.uint32) & mask).view(f32)
    lo = (r - mid).astype(f32)
    assert np.array_equal((lo.view(np.uint32) & np.uint32(0xFFFF)), np.zeros_like(lo.view(np.uint32)))  # lo needs no rounding
    total = hi.astype(np.float64) + mid.astype(np.float64) + lo.astype(np.float64)
    assert np.array_equal(total, x.astype(np.float64))
    # magnitudes: |mid| <= 2^-7 |x|, |lo| <= 2^-15 |x| (truncation), so the three dropped products of a 6-product contraction are
    # below 2^-22 of the leading one
    nz = x != 0
    assert (np.abs(mid[nz]) <= np.abs(x[nz]) * 2.0 ** -7).all() and (np.abs(lo[nz]) <= np.abs(x[nz]) * 2.0 ** -15).all()
