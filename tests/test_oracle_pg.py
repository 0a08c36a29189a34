"""The oracle's own stochastic primitives (CPU, no GPU): Philox4x32-10 against the Random123 known-answer vectors,
the closed-form CDF of PG(1, psi) against itself (two independent series) and against the analytic moments, and the C
Devroye sampler (oracle/pg_devroye.c, stand-in for pypolyagamma.pgdrawvpar, regression.py:501-508) against that CDF.
PARITY UNPINNED for the draws themselves (no golden vector in the reference, pypolyagamma absent): what is pinned is
the distribution."""
import ctypes

import numpy as np
import pytest
from scipy import stats

from oracle import pyglm_oracle as O

PSIS = [0.0, 0.5, 2.0, 5.0, 12.0]


def _philox(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    out = (ctypes.c_uint32 * 4)()
    O.pg_lib().philox4x32_10(c, k, out)
    return [int(v) for v in out]


def test_philox4x32_10_random123_known_answers():
    """kat_vectors of the Random123 distribution (philox4x32, 10 rounds): zeros, all ones, digits of pi."""
    assert _philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert _philox([f, f, f, f], [f, f]) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_stream_layout():
    """Stream convention shared with csrc/philox.cuh: key = seed (lo, hi), counter = (elem lo, elem hi, call, block#);
    a uniform takes two consecutive 32-bit words a, b: ((a >> 5) * 2^26 + (b >> 6)) / 2^53, 53 random bits in [0, 1)."""
    seed, call, elem = (7 << 32) | 5, 9, (3 << 32) | 11
    u = O.philox_uniforms(seed, call, elem, 6)
    w = []
    for blk in range(3):
        w += _philox([11, 3, call, blk], [5, 7])
    w = np.array(w, dtype=np.uint64)
    ref = ((w[0::2] >> np.uint64(5)).astype(np.float64) * 67108864.0 + (w[1::2] >> np.uint64(6)).astype(np.float64)) \
        / 9007199254740992.0
    np.testing.assert_array_equal(u, ref)
    assert not np.array_equal(u, O.philox_uniforms(seed, call + 1, elem, 6))
    assert not np.array_equal(u, O.philox_uniforms(seed, call, elem + 1, 6))


@pytest.mark.parametrize("z", [0.0, 0.25, 1.0, 2.5, 6.0])
def test_pg_cdf_two_series_agree(z):
    x = np.array([0.03, 0.1, 0.3, 0.64, 1.0, 2.0, 5.0])
    left, right = O.jstar_cdf(x, z, "left"), O.jstar_cdf(x, z, "right")
    np.testing.assert_allclose(left, right, rtol=0, atol=1e-13)
    assert np.all(np.diff(left) > 0) and left[0] > 0 and left[-1] <= 1.0


@pytest.mark.parametrize("psi", PSIS)
def test_pg_cdf_reproduces_the_analytic_moments(psi):
    """E[w] = int (1 - F), E[w^2] = int 2 w (1 - F): tanh(psi/2)/(2 psi) and the variance of SURVEY 8(c)."""
    w = np.linspace(0.0, 6.0, 600001)
    surv = 1.0 - O.pg1_cdf(w, psi)
    m1 = np.trapezoid(surv, w)
    m2 = np.trapezoid(2.0 * w * surv, w)
    assert m1 == pytest.approx(float(O.pg1_mean(np.array(psi))), rel=2e-6)
    assert m2 - m1 * m1 == pytest.approx(float(O.pg1_var(np.array(psi))), rel=2e-5)


@pytest.mark.parametrize("rng_kind", [0, 1])
@pytest.mark.parametrize("psi", PSIS)
def test_oracle_sampler_follows_the_pg_law(psi, rng_kind):
    n = 200000
    x = O.pg1_draw(np.full(n, psi), seed=11 + int(10 * psi), call_id=2, rng_kind=rng_kind, nthreads=2)
    assert np.all(x > 0)
    ks = stats.kstest(x, lambda v: O.pg1_cdf(v, psi))
    assert ks.pvalue > 1e-3, ks
    m, v = float(O.pg1_mean(np.array(psi))), float(O.pg1_var(np.array(psi)))
    assert abs(x.mean() - m) < 5 * np.sqrt(v / n)
    assert abs(x.var() - v) < 0.03 * v
    # symmetric in psi (the sampler takes |psi|)
    y = O.pg1_draw(np.full(1000, -psi), seed=11 + int(10 * psi), call_id=2, rng_kind=0)
    np.testing.assert_array_equal(y, O.pg1_draw(np.full(1000, psi), seed=11 + int(10 * psi), call_id=2, rng_kind=0))


def test_oracle_sampler_mixed_psi_probability_integral_transform():
    """psi_t ~ N(-2, 1) (the benchmark's regime): F(omega_t; psi_t) must be uniform."""
    rng = np.random.default_rng(5)
    psi = np.round(rng.standard_normal(400) - 2.0, 2)
    draws = O.pg1_draw(np.repeat(psi, 100), seed=3, call_id=1)
    u = np.concatenate([O.pg1_cdf(draws[i * 100:(i + 1) * 100], p) for i, p in enumerate(psi)])
    assert stats.kstest(u, "uniform").pvalue > 1e-3
