"""CPU: the compile-time layout of the register-resident CTA kernel (cyrk_b200/csrc/cta_core_params.h) -- which working
vector lives in registers / shared memory / the L2 slab for every (method, components per thread) -- checked by compiling the
header's constexpr functions with the host compiler: shared memory fits the 227 KB of an SM, every stage derivative has
exactly one home, the staging vector exists, and the choice of C covers n_y up to 12 288 without gaps."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cstdio>
#include "cta_core_params.h"
using namespace cyrk;
int main() {
    const int Cs[] = {1, 2, 4, 8, 12, 16, 20, 24};
    for (int method = 3; method <= 5; method++)
        for (int C : Cs) {
            const int nkd = cta2_nkd(method);
            std::printf("%d %d %d %d %d %d %d %d %d %d\n", method, C, nkd, cta2_nreg(nkd, C), cta2_nsm(nkd, C), cta2_ngl(nkd, C),
                        cta2_smem_vectors(nkd, C), cta2_slab_vectors(nkd, C), cta2_aux_global(C) ? 1 : 0, cta2_threads(C));
        }
    for (unsigned long long n = 1; n <= 13000; n++) std::printf("C %llu %d\n", n, cta2_choose_c(n));
    return 0;
}
'''


@pytest.fixture(scope="module")
def layout(tmp_path_factory):
    d = tmp_path_factory.mktemp("cta_layout")
    src = d / "layout.cpp"
    src.write_text(SRC)
    exe = d / "layout"
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "cyrk_b200", "csrc"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()
    rows = [tuple(int(x) for x in l.split()) for l in out if not l.startswith("C ")]
    choose = {int(l.split()[1]): int(l.split()[2]) for l in out if l.startswith("C ")}
    return rows, choose


def test_every_stage_derivative_has_one_home_and_shared_memory_fits(layout):
    rows, _ = layout
    assert len(rows) == 3 * 8
    for method, C, nkd, nreg, nsm, ngl, smem_vec, slab_vec, aux_global, threads in rows:
        assert nkd == {3: 4, 4: 7, 5: 16}[method]
        assert nreg >= 0 and nsm >= 0 and ngl >= 0 and nreg + nsm + ngl == nkd, (method, C)
        V = C * threads
        assert smem_vec * V * 8 + 2048 <= 227 * 1024, (method, C)            # dynamic + static shared memory of one CTA
        assert (1 + nreg) * C * 2 <= 96, (method, C)                        # registers of state per thread (of 128)
        # two stage-input buffers always; aux in shared memory unless a third vector does not fit
        assert smem_vec >= (2 if aux_global else 3) + nsm
        assert smem_vec - ((2 if aux_global else 3) + nsm) == (1 if max(nreg, 1) >= nkd else 0)    # staging vector of its own
        # slab: the memory-resident derivatives that are not in shared memory, one vector for root searches, aux if needed
        assert slab_vec == max(ngl, 1) + 1 + aux_global
        if max(nreg, 1) < nkd:
            assert nsm + ngl >= 1                                            # K_{max(NREG,1)} lives in memory: it is the staging vector
        if C <= 8 and method != 5:
            assert ngl == 0, "RK23 / RK45 up to n_y = 4096 stay on chip"


def test_choice_of_components_per_thread(layout):
    _, choose = layout
    prev = 1
    for n in range(1, 12289):
        C = choose[n]
        assert C in (1, 2, 4, 8, 12, 16, 20, 24) and C >= prev
        assert n <= C * 512 and (C == 1 or n > {2: 512, 4: 1024, 8: 2048, 12: 4096, 16: 6144, 20: 8192, 24: 10240}[C])
        prev = C
    assert choose[12289] == 0 and choose[13000] == 0          # beyond: the shared-memory/slab kernel
    assert choose[4096] == 8 and choose[10000] == 20
