"""SIC-CDF-1 (oracle/cdf_exact.c) against scipy float64, plus structural properties of the tables."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st
from scipy import special as sp

from oracle import clib
from oracle import numpy_ref as R


def test_tcdf_matches_scipy():
    rng = np.random.default_rng(0)
    nu = rng.uniform(1.1, 100, 4000)
    t = rng.standard_t(3, 4000) * rng.choice([0.01, 1, 10, 1000], 4000)
    mine, ref = clib.tcdf(t, nu), sp.stdtr(nu, t)
    assert np.max(np.abs(mine - ref) / ref) < 5e-12
    assert clib.tcdf(0.0, 7.0) == 0.5
    assert clib.tcdf(np.inf, 3.0) == 1.0 and clib.tcdf(-np.inf, 3.0) == 0.0


def test_tcdf_symmetry_and_monotone():
    nu = 5.5
    t = np.linspace(-30, 30, 601)
    F = clib.tcdf(t, nu)
    assert (np.diff(F) > 0).all()
    np.testing.assert_allclose(F + F[::-1], 1.0, atol=1e-14)


def test_ncdf_matches_scipy():
    t = np.linspace(-35, 9, 3001)
    mine, ref = clib.ncdf(t), sp.ndtr(t)
    assert np.max(np.abs(mine - ref) / ref) < 1e-12


@settings(max_examples=60, deadline=None)
@given(st.floats(1e-3, 50.0), st.floats(2.0, 100.0), st.integers(-40, 20), st.integers(1, 60))
def test_table_properties(sigma, nu, mn, L):
    mx = mn + L - 1
    for kind in ("studentt", "gaussian"):
        t = clib.build_tables(kind, [sigma], [nu], [0], [mn], [mx])[0].astype(np.int64)
        assert t.shape == (L + 1,)
        assert t[0] == 0 and t[-1] == 65535
        assert (np.diff(t) >= 0).all()


def test_table_matches_numpy_flow():
    """C table row == numpy statement of the same flow (fp32 CDF values -> pmf -> spec'd pmf_to_uint16_cdf)."""
    rng = np.random.default_rng(1)
    for _ in range(20):
        sigma, nu = np.float32(np.exp(rng.normal())), np.float32(rng.uniform(2, 100))
        mn, L = int(rng.integers(-30, 5)), int(rng.integers(2, 50))
        edges = (np.arange(mn, mn + L + 1, dtype=np.float32) - np.float32(0.5))
        Fv = clib.tcdf((edges / sigma).astype(np.float64), float(nu)).astype(np.float32)
        pmf = np.maximum(Fv[1:] - Fv[:-1], np.float32(1e-12))
        S = np.float32(pmf.astype(np.float64).cumsum()[-1])
        ref = R.pmf_to_uint16_cdf_spec((pmf / S).astype(np.float32))
        got = clib.build_tables("studentt", [sigma], [nu], [0], [mn], [mn + L - 1])[0]
        assert np.array_equal(got, ref)


def test_bin_probability_consistent_with_tables():
    """L2 oracle (same-side survival differences) agrees with CDF differences where those are well conditioned."""
    x = np.arange(-6, 7, dtype=np.float64)
    p = R.studentt_bin_prob_f64(x, 1.3, 4.0)
    ref = sp.stdtr(4.0, (x + 0.5) / 1.3) - sp.stdtr(4.0, (x - 0.5) / 1.3)
    np.testing.assert_allclose(p, ref, rtol=1e-10)
    assert abs(R.studentt_bin_prob_f64(np.arange(-4000, 4001, dtype=np.float64), 1.3, 4.0).sum() - 1) < 1e-8
