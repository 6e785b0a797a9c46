"""The two-trajectories-per-thread variant of the final-state kernels (rk_core.cuh: rk_thread_body2 / attempt_pair, knob
CYRK_SLOTS2, off by default because it measured 15-20 % slower: profiles/r02_notes.md section 9) gives the one-slot kernel's
results bit for bit: per trajectory it runs the same instructions on the same inputs.  Built here through NVRTC with the
knob as an extra definition (CYRK_B200_NVRTC_DEFINES), so the variant stays buildable and checked."""
import re

import numpy as np
import pytest

import cyrk_b200 as c
from test_user_rhs import LV_SRC

FIELDS = ("status", "steps_taken", "n_rejected", "nfev", "size", "t_last", "y_last")


def _ctas_per_sm(err):
    m = re.findall(r"(\d+) CTAs/SM", err)
    return int(m[-1]) if m else None


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["RK23", "RK45"])
def test_two_slot_kernel_equals_one_slot_bitwise(method, monkeypatch, capfd):
    # two registrations of the same function (the trailing comment makes them different sources: separate kernel caches)
    one = c.UserRHS(LV_SRC + "// one trajectory per thread\n", "my_lv", n_y=2, args_bytes=32)
    two = c.UserRHS(LV_SRC + "// two trajectories per thread\n", "my_lv", n_y=2, args_bytes=32)
    assert one.rhs_id != two.rhs_id
    monkeypatch.setenv("CYRK_B200_DEBUG", "1")
    rng = np.random.default_rng(11)
    occ = {}
    # 1 / 77: slots that never hold a trajectory; 4099: partially filled warps; 300 000: refills, start-up kernel
    for B in (1, 77, 4099, 300_000):
        y0 = np.stack([rng.uniform(5, 15, B), rng.uniform(2.5, 7.5, B)], 1)
        # backward runs among the trajectories
        ts = np.tile([0.0, 10.0], (B, 1))
        ts[::7] = [10.0, 0.0]
        kw = dict(method=method, args=[1.5, 1, 3, 1], rtol=1e-7, atol=1e-9, store="final")
        monkeypatch.delenv("CYRK_B200_NVRTC_DEFINES", raising=False)
        capfd.readouterr()
        a = c.cysolve_ivp_batch(one, ts, y0, **kw)
        occ["one"] = _ctas_per_sm(capfd.readouterr().err) or occ.get("one")
        monkeypatch.setenv("CYRK_B200_NVRTC_DEFINES", "-DCYRK_SLOTS2=1")
        b = c.cysolve_ivp_batch(two, ts, y0, **kw)
        occ["two"] = _ctas_per_sm(capfd.readouterr().err) or occ.get("two")
        assert (np.asarray(a.status) == 1).all()
        for f in FIELDS:
            assert np.array_equal(np.asarray(getattr(a, f)), np.asarray(getattr(b, f)), equal_nan=True), (B, f)
    # the variant really was the two-slot kernel: twice the state per thread, fewer resident CTAs
    assert occ["one"] is not None and occ["two"] is not None and occ["two"] < occ["one"], occ
