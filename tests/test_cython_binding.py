"""CPU: the Cython declarations (include/cyrk_b200.pxd) compile against include/cyrk_b200.h, link against the built
library and can be called `nogil` -- the binding a CyRK maintainer would add next to cy/cysolver_api.pxd
(INTEGRATION.md section 2).  The C compiler sees the header's prototypes and Cython's declarations together, so a
signature that drifts apart fails the build; every function the .pxd declares is referenced once."""
import importlib.util
import os
import re
import subprocess
import sys
import sysconfig

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PYX_HEAD = '''# cython: language_level=3
from libc.stddef cimport size_t
cimport cyrk_b200 as cb

def version():
    cdef int v
    with nogil:
        v = cb.cyrk_b200_version()
    return v

def device_count():
    cdef int n
    with nogil:
        n = cb.cyrk_b200_device_count()
    return n

def builtin_id(bytes name):
    cdef const char* s = name
    cdef int r
    with nogil:
        r = cb.cyrk_b200_rhs_builtin(s)
    return r

def message(int code):
    cdef char buf[256]
    cb.cyrk_b200_status_message(code, buf, 256)
    return bytes(buf).split(b"\\0")[0].decode()

def constants():
    return (cb.CYRK_METHOD_RK23, cb.CYRK_METHOD_RK45, cb.CYRK_METHOD_DOP853, cb.CYRK_RHS_LOTKAVOLTERRA, cb.CYRK_STORE_FINAL)

def addresses():
    """Takes the address of every declared function: each must exist with a compatible prototype."""
    return [
'''


def test_pxd_compiles_links_and_calls(tmp_path):
    try:
        import Cython  # noqa: F401
    except ImportError:
        pytest.skip("Cython is not installed")
    lib = os.path.join(ROOT, "cyrk_b200", "libcyrk_b200.so")
    assert os.path.exists(lib), "build the library first (__graft_entry__.build())"
    pxd = open(os.path.join(ROOT, "include", "cyrk_b200.pxd")).read()
    funcs = re.findall(r"\b(cyrk_b200_\w+)\s*\(", pxd)
    hdr = open(os.path.join(ROOT, "include", "cyrk_b200.h")).read()
    declared_in_header = set(re.findall(r"\b(cyrk_b200_\w+)\s*\(", hdr))
    assert set(funcs) == declared_in_header, set(funcs) ^ declared_in_header     # the .pxd covers the whole header
    body = PYX_HEAD + "".join(f"        <size_t>&cb.{f},\n" for f in funcs) + "    ]\n"
    pyx = tmp_path / "cyrk_b200_cy_probe.pyx"
    pyx.write_text(body)
    inc = os.path.join(ROOT, "include")
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    subprocess.run([sys.executable, "-m", "cython", "-3", "-I", inc, str(pyx), "-o", str(tmp_path / "probe.c")], check=True, env=env)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = tmp_path / ("cyrk_b200_cy_probe" + ext)
    subprocess.run(["/usr/bin/gcc", "-shared", "-fPIC", "-O1", "-I", inc, "-I", sysconfig.get_paths()["include"],
                    str(tmp_path / "probe.c"), "-o", str(so), lib, f"-Wl,-rpath,{os.path.dirname(lib)}"], check=True, env=env)
    spec = importlib.util.spec_from_file_location("cyrk_b200_cy_probe", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.version() >= 100
    assert mod.device_count() >= 0
    assert mod.builtin_id(b"lotkavolterra") == mod.constants()[3]
    assert mod.constants()[:3] == (3, 4, 5)                      # ODEMethod ints of the reference (cysolver.hpp:34-41)
    assert "success" in mod.message(1).lower() or len(mod.message(1)) > 0
    assert all(a != 0 for a in mod.addresses()) and len(mod.addresses()) == len(funcs)
