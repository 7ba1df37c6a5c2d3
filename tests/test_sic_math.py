"""Accuracy of the bit-reproducible elementary functions the kernels and the oracle share (csrc/sic_math.h, through
the oracle's C shim): < 1 ulp against mpmath over the argument ranges the constitutive laws reach, C99 special cases of
pow, and the exact small-integer fast paths."""
import numpy as np
import pytest

from oracle import constitutive as oc

mp = pytest.importorskip("mpmath")
mp.mp.prec = 200


def ulp_error(got, exact_mpf):
    exact = float(exact_mpf)
    if exact == 0.0 or not np.isfinite(exact):
        return 0.0 if got == exact else np.inf
    return float(abs(mp.mpf(float(got)) - exact_mpf) / mp.mpf(float(np.spacing(abs(exact)))))


def test_exp_log_within_one_ulp():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-40.0, 40.0, 400), rng.uniform(-1e-3, 1e-3, 100), [-700.0, 700.0, 0.0, 1.0, -1.0]])
    got = oc.f_exp(x)
    assert max(ulp_error(g, mp.exp(mp.mpf(float(v)))) for g, v in zip(got, x)) < 1.0
    y = np.concatenate([10.0 ** rng.uniform(-12, 12, 400), 1.0 + rng.uniform(-1e-6, 1e-6, 100), [1.0, 2.0, 0.5, 5e-324]])
    import ctypes
    out = np.empty_like(y)
    oc._lib().sic_log_array(y.ctypes.data_as(oc._P), out.ctypes.data_as(oc._P), ctypes.c_long(y.size))
    assert max(ulp_error(g, mp.log(mp.mpf(float(v)))) for g, v in zip(out, y)) < 1.0


def test_pow_within_one_ulp_over_the_laws_ranges():
    rng = np.random.default_rng(1)
    # von Mises stresses in Pa / MPa, J2, hardening variables ... with the exponents of the element library
    x = np.concatenate([10.0 ** rng.uniform(3, 8, 300), 10.0 ** rng.uniform(-6, 2, 300)])
    for y in (1.5, 2.5, -0.5, 3.0, 3.99, 4.99, 0.8275682807874163, 1.0 / 0.8275682807874163, 7.0, -2.0):
        got = oc.f_pow(x, np.full_like(x, y))
        err = max(ulp_error(g, mp.power(mp.mpf(float(v)), mp.mpf(y))) for g, v in zip(got, x))
        assert err < 1.0, (y, err)


def test_pow_exact_small_integer_exponents_and_special_cases():
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.standard_normal(200) * 1e7, [0.0, -0.0, np.inf, -np.inf, 1.0, -3.0]])
    assert np.array_equal(oc.f_pow(x, np.full_like(x, 2.0)), x * x)          # correctly rounded for every x
    assert np.array_equal(oc.f_pow(x, np.full_like(x, 1.0)), x)
    assert np.array_equal(oc.f_pow(x, np.zeros_like(x)), np.ones_like(x))
    assert np.isnan(oc.f_pow(np.array([np.nan]), np.array([2.0]))[0])
    with np.errstate(all="ignore"):
        ref = np.power(np.array([-8.0, -8.0, 0.0, 0.0, np.inf, 2.0, 0.5]), np.array([3.0, 0.5, -1.0, 2.5, -1.0, np.inf, np.inf]))
    got = oc.f_pow(np.array([-8.0, -8.0, 0.0, 0.0, np.inf, 2.0, 0.5]), np.array([3.0, 0.5, -1.0, 2.5, -1.0, np.inf, np.inf]))
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.allclose(got[~np.isnan(ref)], ref[~np.isnan(ref)], rtol=1e-15)
