"""Generates tests/golden/golden_v1.npz by running the UNMODIFIED reference (oracle/_ref/libcyrk_ref.so,
compiled from /root/reference by oracle/Makefile) on the cases of cases.py.

    python tests/golden/make_golden.py        # only where oracle/_ref has been built

The fixture is committed; the GPU box never needs /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from oracle import oracle as o   # noqa: E402
from cases import cases          # noqa: E402


def main():
    assert o.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for name, kw in cases().items():
        r = o.solve(impl="ref", **kw)
        out[f"{name}/scalars"] = np.array([r.status, r.size, r.steps_taken, r.num_interpolates, r.nfev,
                                           int(r.event_terminated),
                                           -1 if r.event_terminate_index == o.SIZE_MAX else r.event_terminate_index],
                                          dtype=np.int64)
        out[f"{name}/t"] = r.t
        out[f"{name}/y"] = r.y
        for e, (te, ye) in enumerate(zip(r.t_events, r.y_events)):
            out[f"{name}/t_events{e}"] = te
            out[f"{name}/y_events{e}"] = ye
        if r.dense is not None:
            out[f"{name}/dense"] = r.dense
            out[f"{name}/dense_code"] = np.array([r.dense_code])
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(cases())} cases, {os.path.getsize(path)/1e3:.0f} kB")


if __name__ == "__main__":
    main()
