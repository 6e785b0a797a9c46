"""Spring-chain systems with 4 <= n_y <= 32 for the thread-per-trajectory path: m masses, y = (x_1..x_m, v_1..v_m),
dx_i = v_i, dv_i = -k (2 x_i - x_{i-1} - x_{i+1}) - c v_i.  CUDA source for NVRTC and the same arithmetic (same order,
no library functions) as a ctypes callback for the CPU oracle."""
import ctypes as C

import numpy as np


def chain_source(m: int) -> str:
    return f"""
#define CHAIN_M {m}
__device__ void chain{2 * m}(double* dy, double t, const double* y, const char* args) {{
    const double* a = (const double*)args;
    const double k = a[0], c = a[1];
#pragma unroll
    for (int i = 0; i < CHAIN_M; i++) {{
        const double left = i > 0 ? y[i - 1] : 0.0, right = i + 1 < CHAIN_M ? y[i + 1] : 0.0;
        dy[i] = y[CHAIN_M + i];
        dy[CHAIN_M + i] = -k * (2.0 * y[i] - left - right) - c * y[CHAIN_M + i];
    }}
}}
"""


def chain_callback(oracle, m: int):
    n = 2 * m

    @oracle.DIFFEQ_FN
    def cb(dy, t, y, args, pre):
        yy = np.ctypeslib.as_array(y, shape=(n,))
        a = C.cast(args, C.POINTER(C.c_double))
        k, c = a[0], a[1]
        x, v = yy[:m], yy[m:]
        left = np.concatenate([[0.0], x[:-1]])
        right = np.concatenate([x[1:], [0.0]])
        out = np.concatenate([v, -k * (2.0 * x - left - right) - c * v])
        C.memmove(dy, out.ctypes.data, 8 * n)
    return cb


def chain_inputs(m: int, B: int, seed: int):
    rng = np.random.default_rng(seed)
    y0 = np.concatenate([rng.uniform(-1.0, 1.0, (B, m)), rng.uniform(-0.5, 0.5, (B, m))], 1)
    args = np.stack([rng.uniform(0.5, 2.0, B), rng.uniform(0.0, 0.2, B)], 1)
    return np.ascontiguousarray(y0), np.ascontiguousarray(args)
