"""Helper of tests/test_gpu_advect_variants.py: runs the advect stage and two whole steps on
seeded inputs with whatever UBGL_ADVECT_VARIANT / UBGL_ADVECT_DIV the environment selects (fixed
per process) and dumps the velocity fields, p and f.
    python tests/advect_dump.py OUT.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import cases  # noqa: E402

SIZES = [(41, 33), (70, 40), (77, 41), (130, 97), (258, 131), (1090, 436)]


def main(out):
    import ubootgl_b200 as u
    from ubootgl_b200 import capi
    res = {}
    for W, H in SIZES:
        c = cases.sim_case(W, H, seed=W * 7 + H)
        for dt_scale, tag in ((1.0, "cfl1"), (9.0, "cfl9")):
            s = u.Simulation(c["flag"])
            s.set(capi.VX, c["vx"]); s.set(capi.VY, c["vy"])
            s.set(capi.VXB, c["vx"][::-1].copy()); s.set(capi.VYB, c["vy"][::-1].copy())
            dt = float(np.float32(0.8) / np.float32(W - 1)) * dt_scale
            s.stage(capi.ST_ADVECT, dt)
            res[f"{W}x{H}_{tag}_vx"] = s.get(capi.VX)
            res[f"{W}x{H}_{tag}_vy"] = s.get(capi.VY)
            s.step(dt); s.step(dt)
            res[f"{W}x{H}_{tag}_vx2"] = s.get(capi.VX)
            res[f"{W}x{H}_{tag}_vy2"] = s.get(capi.VY)
            res[f"{W}x{H}_{tag}_p2"] = s.get(capi.P)
            res[f"{W}x{H}_{tag}_f2"] = s.get(capi.F)
    np.savez(out, **res)


if __name__ == "__main__":
    main(sys.argv[1])
