"""CPU: run-time compiled kernels (NVRTC needs no GPU) are cached on disk, keyed by source + options + NVRTC version.
SURVEY.md section 7 step 8 / VERDICT r1 item 9: "cubins cached by source hash"."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROBE = textwrap.dedent("""
    import ctypes as C, sys
    sys.path.insert(0, %r)
    import cyrk_b200 as c
    L = c.lib()
    src = '''
    __device__ void osc(double* dy, double t, const double* y, const char* args) {
        const double* a = (const double*)args;
        dy[0] = y[1]; dy[1] = -a[0] * y[0];
    }'''
    u = c.UserRHS(src, "osc", n_y=2, args_bytes=8)                     # validates by compiling one kernel
    assert L.cyrk_b200_jit_compile(u.rhs_id, 3, 1, 0) == 0             # RK23, rows variant
    assert L.cyrk_b200_jit_compile(5, 4, 0, 8) == 0, c.last_error()    # built-in Lotka-Volterra, per-trajectory tolerances
    assert L.cyrk_b200_jit_compile(5, 4, 0, 8 | 1) == 0                # ... strict FP: a different kernel
    a, b = C.c_ulonglong(), C.c_ulonglong()
    L.cyrk_b200_jit_cache_stats(C.byref(a), C.byref(b))
    print(a.value, b.value)
""") % ROOT


def _run(cache_dir, extra_env=None):
    env = dict(os.environ, CYRK_B200_CACHE_DIR=str(cache_dir))
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, "-c", PROBE], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return tuple(int(x) for x in out.stdout.split()[-2:])


def test_second_process_finds_every_kernel_on_disk(tmp_path):
    d = tmp_path / "cache"
    compiled, hits = _run(d)
    assert compiled == 4 and hits == 0
    files = sorted(os.listdir(d))
    assert len(files) == 4 and all(f.endswith(".cubin") and len(f) == 32 + 6 for f in files)
    compiled, hits = _run(d)
    assert compiled == 0 and hits == 4
    # a truncated file is not trusted: it is recompiled and replaced
    victim = d / files[0]
    victim.write_bytes(victim.read_bytes()[:10])
    compiled, hits = _run(d)
    assert compiled == 1 and hits == 3
    # and the cache can be switched off
    compiled, hits = _run(d, {"CYRK_B200_CACHE": "0"})
    assert compiled == 4 and hits == 0
