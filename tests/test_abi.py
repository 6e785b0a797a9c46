"""CPU: the C-ABI library loads, exports every symbol include/cyrk_b200.h declares, and -- with no GPU -- the
product path fails loudly instead of falling back to anything."""
import os
import re

import numpy as np
import pytest

import cyrk_b200 as c

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cyrk_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cyrk_b200_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 35
    lib = c.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cyrk_b200.h but not exported"
        assert n in c.SIGNATURES, f"{n} has no ctypes signature in cyrk_b200/_lib.py"
    assert sorted(c.SIGNATURES) == names


def test_builtin_ids_and_info():
    lib = c.lib()
    assert lib.cyrk_b200_version() >= 100
    import ctypes as C
    for name, rid, ny, nx in [("baseline", 0, 2, 0), ("accuracy", 1, 2, 0), ("extraoutput", 2, 2, 2), ("lorenz", 3, 3, 0),
                              ("lorenz_extraoutput", 4, 3, 3), ("lotkavolterra", 5, 2, 0), ("pendulum", 6, 2, 0),
                              ("pendulum_drag", 11, 2, 1), ("dense_kat", 12, 2, 2)]:
        assert lib.cyrk_b200_rhs_builtin(name.encode()) == rid
        a, b, d = C.c_size_t(), C.c_size_t(), C.c_size_t()
        assert lib.cyrk_b200_rhs_info(rid, C.byref(a), C.byref(b), C.byref(d)) == 0
        assert (a.value, b.value) == (ny, nx)
    assert lib.cyrk_b200_rhs_builtin(b"nope") == -1


def test_status_messages_match_reference_table():
    # CyRK/cy/c_common.hpp:84-177
    assert c.status_message(1) == "Integration completed without issue."
    assert c.status_message(2).startswith("Integration ended early: An event has been triggered")
    assert c.status_message(-52) == "Maximum number of steps (set by user) exceeded during integration."
    assert c.status_message(-53) == "Maximum number of steps (set by system architecture) exceeded during integration."
    assert c.status_message(-54) == "Error in step size calculation: Required step size is less than spacing between numbers."
    assert c.status_message(-80) == "User-provided initial step size must be a positive number."
    assert c.status_message(12345) == "UNKNOWN_ERROR"


def test_every_status_code_the_library_can_emit_has_a_message():
    """cyrk_defs.h `enum Status` is the set of codes device code writes; each one must come back with a text (the
    reference's table has no entry for -10 / -30 / -31, which this library does emit)."""
    src = open(os.path.join(ROOT, "cyrk_b200", "csrc", "cyrk_defs.h")).read()
    body = src[src.index("enum Status"):]
    body = body[:body.index("};")]
    codes = [int(v) for v in re.findall(r"=\s*(-?\d+)", body)]
    assert len(codes) >= 20 and -10 in codes and -30 in codes and -31 in codes
    for code in codes + [-60]:
        assert c.status_message(code) != "UNKNOWN_ERROR", code


def test_non_finite_or_unordered_times_are_rejected_before_any_launch():
    """A NaN in t_span / t_eval used to reach the kernels (out-of-bounds t_eval walk); validation runs before the
    device is touched, so this also holds on a box without a GPU."""
    y0 = np.ones((3, 2))
    kw = dict(method="RK45", args=[1.5, 1, 3, 1])
    with pytest.raises(c.CyrkB200Error, match="not finite"):
        c.cysolve_ivp_batch("lotkavolterra", (float("nan"), 1.0), y0, **kw)
    with pytest.raises(c.CyrkB200Error, match="not finite"):
        c.cysolve_ivp_batch("lotkavolterra", (0.0, float("inf")), y0, **kw)
    with pytest.raises(c.CyrkB200Error, match="not finite"):
        c.cysolve_ivp_batch("lotkavolterra", (0.0, 1.0), y0, t_eval=[0.0, float("nan"), 1.0], **kw)
    with pytest.raises(c.CyrkB200Error, match="ordered in the direction"):
        c.cysolve_ivp_batch("lotkavolterra", (0.0, 1.0), y0, t_eval=[0.0, 0.5, 0.25], **kw)
    with pytest.raises(c.CyrkB200Error, match="ordered in the direction"):
        c.cysolve_ivp_batch("lotkavolterra", (1.0, 0.0), y0, t_eval=[0.0, 0.5, 1.0], **kw)
    with pytest.raises(c.CyrkB200Error, match="same direction"):
        c.cysolve_ivp_batch("lotkavolterra", np.array([[0.0, 1.0], [1.0, 0.0], [0.0, 1.0]]), y0, t_eval=[0.0, 0.5], **kw)
    with pytest.raises(c.CyrkB200Error, match="unknown event id"):
        c.cysolve_ivp_batch("largey", (0.0, 1.0), np.ones((2, 64)), args=[-0.5, 64.0], events=[c.Event(99)])


def test_python_layer_errors_match_reference():
    """pysolver.pyx:83,257,262 -- raised before anything touches the device."""
    y0 = np.ones((3, 2))
    with pytest.raises(NotImplementedError):
        c.cysolve_ivp_batch("lotkavolterra", (0, 1), y0, method="FakeIntegrationMethod", args=[1, 1, 1, 1])
    with pytest.raises(AttributeError):
        c.cysolve_ivp_batch("lotkavolterra", (0, 1), y0, method="RK23", first_step=-10.0, args=[1, 1, 1, 1])
    with pytest.raises(AttributeError):
        c.cysolve_ivp_batch("lotkavolterra", (0, 1), y0, method="RK23", max_step=-10.0, args=[1, 1, 1, 1])
    with pytest.raises(TypeError):
        c.cysolve_ivp_batch(lambda t, y: y, (0, 1), y0)
    with pytest.raises(NotImplementedError):
        c.cysolve_ivp_batch("not_a_builtin", (0, 1), y0)


def test_no_gpu_means_loud_failure():
    if c.lib().cyrk_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(c.CyrkB200Error, match="no CPU fallback"):
        c.cysolve_ivp_batch("lotkavolterra", (0, 10), np.ones((4, 2)), args=[1.5, 1, 3, 1])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cyrk_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt and "libcyrk_oracle" not in txt \
                    and "libcyrk_ref" not in txt, f
