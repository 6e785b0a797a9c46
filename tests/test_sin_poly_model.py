"""CPU: the arithmetic of psin_pi (cyrk_b200/csrc/rk_core.cuh: reduction by pi in three FMA steps, one degree-7 polynomial in
r^2, sign from the quotient's parity) restated in C with the SAME constants -- read from the header, so the two cannot drift
apart -- and checked against long-double sinl: <= 2 ulp on |x| <= 1e5, multiples of pi/2 and their neighbours included.
(The GPU test tests/test_gpu_math.py checks the device function itself; this one pins the coefficients without a GPU.)"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_SRC = r'''
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
static const double C[8] = {%(coefs)s};
static double psin_pi(double x) {
    const double MAGIC = 6755399441055744.0;
    double qd = fma(x, %(inv_pi)s, MAGIC);
    uint64_t b; memcpy(&b, &qd, 8); const int q = (int)(uint32_t)b;
    qd -= MAGIC;
    double r = fma(qd, %(pi_hi)s, x);
    r = fma(qd, %(pi_mid)s, r);
    r = fma(qd, %(pi_lo)s, r);
    const double z = r * r;
    double p = C[7];
    for (int k = 6; k >= 0; k--) p = fma(p, z, C[k]);
    const double res = fma(r * z, p, r);
    return (q & 1) ? -res : res;
}
int main(void) {
    double mx = 0, worst = 0; srand48(1);
    for (long i = 0; i < 4000000; i++) {
        double x; const int m = (int)(i %% 5);
        if (m == 0) x = (drand48() * 2 - 1) * M_PI; else if (m == 1) x = (drand48() * 2 - 1) * 100; else if (m == 2) x = (drand48() * 2 - 1) * 1e5;
        else if (m == 3) x = (drand48() * 2 - 1) * 1e-3; else { const long k = lrand48() %% 60000 - 30000; x = k * (M_PI / 2); if (lrand48() & 1) x = nextafter(x, 0); }
        const long double ref = sinl((long double)x);
        const double rd = (double)ref, u = nextafter(fabs(rd), INFINITY) - fabs(rd);
        if (u == 0) continue;
        const double e = (double)(fabsl((long double)psin_pi(x) - ref) / u);
        if (e > mx) { mx = e; worst = x; }
    }
    printf("%%.4f %%.17g\n", mx, worst);
    return 0;
}
'''


def _header_constants():
    src = open(os.path.join(ROOT, "cyrk_b200", "csrc", "rk_core.cuh")).read()
    # g_sinpi = {{ c[0..7] }, { 1/pi, -pi (three pieces), magic }}: polynomial and reduction constants live in one table
    m = re.search(r"g_sinpi = \{\{(.*?)\},\s*\{(.*?)\}\};", src, re.S)
    tab, red = m.group(1), [x.strip() for x in m.group(2).split(",")]
    assert len(red) == 5 and float(red[4]) == 6755399441055744.0
    body = src[src.index("CYRK_DI double psin_pi(double x)"):]
    body = body[:body.index("\n}\n")]
    # the function uses the table entries in this order
    assert "fma(x, g_sinpi.r[0], MAGIC)" in body and "fma(qd, g_sinpi.r[1], x)" in body
    assert "fma(qd, g_sinpi.r[2], r)" in body and "fma(qd, g_sinpi.r[3], r)" in body
    return dict(coefs=" ".join(tab.split()), inv_pi=red[0], pi_hi=red[1], pi_mid=red[2], pi_lo=red[3])


def test_psin_pi_model_within_two_ulp(tmp_path):
    k = _header_constants()
    assert len(k["coefs"].split(",")) == 8
    src = tmp_path / "sinpi.c"
    src.write_text(C_SRC % k)
    exe = tmp_path / "sinpi"
    # -mfma: fma() must be the hardware instruction for speed; the result is the same correctly rounded value either way
    flags = ["-O2", "-mfma"] if "fma" in open("/proc/cpuinfo").read() else ["-O2"]
    subprocess.run(["gcc", *flags, str(src), "-o", str(exe), "-lm"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert float(out[0]) <= 2.0, f"psin_pi model: {out[0]} ulp at x = {out[1]}"
