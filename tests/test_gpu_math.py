"""GPU: accuracy of the production build's sin/cos for the built-in right-hand sides (rk_core.cuh psin/pcos)
against libm evaluated in extended precision."""
import ctypes as C

import numpy as np
import pytest

import cyrk_b200 as c

pytestmark = pytest.mark.gpu


def _device_sincos(x, mode=0):
    x = np.ascontiguousarray(x, dtype=np.float64)
    s = np.empty_like(x); co = np.empty_like(x)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert c.lib().cyrk_b200_debug_sincos(p(x), p(s), p(co), x.size, mode << 8) == 0
    return s, co


def _ulp_err(got, x, fn):
    ref = fn(x.astype(np.longdouble))
    return np.abs((got.astype(np.longdouble) - ref) / np.spacing(np.abs(ref.astype(np.float64))).astype(np.longdouble)).astype(np.float64)


def test_psin_pcos_within_two_ulp():
    rng = np.random.default_rng(0)
    x = np.concatenate([
        rng.uniform(-np.pi, np.pi, 200000), rng.uniform(-100, 100, 200000), rng.uniform(-1e5, 1e5, 200000),
        rng.uniform(-1e-3, 1e-3, 50000), np.arange(-2000, 2001) * (np.pi / 2), np.arange(-2000, 2001) * (np.pi / 4),
        np.nextafter(np.arange(1, 2001) * (np.pi / 2), 0.0), [0.0, -0.0, 1e-300, -1e-300, 5e-324, 1e5, -1e5]])
    s, co = _device_sincos(x)
    assert _ulp_err(s, x, np.sin).max() <= 2.0
    assert _ulp_err(co, x, np.cos).max() <= 2.0
    z = _device_sincos(np.array([0.0, -0.0]))             # (the sign of a zero result is not preserved: sin(-0.0) = +0.0)
    assert np.array_equal(z[0], [0.0, 0.0]) and np.array_equal(z[1], [1.0, 1.0])


def test_check_free_sin_variants_within_two_ulp():
    """psin_pi (one polynomial after reduction by pi: the CTA kernels' largey functor) and psin_fast, |x| <= 1e5."""
    rng = np.random.default_rng(1)
    x = np.concatenate([
        rng.uniform(-np.pi, np.pi, 200000), rng.uniform(-100, 100, 200000), rng.uniform(-1e5, 1e5, 400000),
        rng.uniform(-1e-3, 1e-3, 50000), np.arange(-30000, 30001) * (np.pi / 2), np.arange(-2000, 2001) * (np.pi / 4),
        np.nextafter(np.arange(1, 30001) * (np.pi / 2), 0.0), [0.0, 1e-300, -1e-300, 5e-324, 1e5, -1e5]])
    s_pi, s_fast = _device_sincos(x, mode=1)
    assert _ulp_err(s_pi, x, np.sin).max() <= 2.0
    assert _ulp_err(s_fast, x, np.sin).max() <= 2.0


def test_psin_pcos_large_and_special_arguments_use_the_library_path():
    x = np.array([1.0000001e5, -3.3e7, 1e22, -1e300, np.inf, -np.inf, np.nan])
    s, co = _device_sincos(x)
    fin = np.isfinite(x)
    assert _ulp_err(s[fin], x[fin], np.sin).max() <= 2.0 and _ulp_err(co[fin], x[fin], np.cos).max() <= 2.0
    assert np.isnan(s[~fin]).all() and np.isnan(co[~fin]).all()
