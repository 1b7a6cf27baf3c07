"""Generates tests/golden/* from the UNMODIFIED reference (oracle/_ref, i.e.
/root/reference/{pressure_solver,simulation,terrain}.cpp compiled where they
lie).  Run in the build container only:  python tests/golden/make_golden.py

The fixtures are small (bit-packed masks, <= 130x97 float fields) and are what
pins the C restatement and the CUDA path on boxes where /root/reference does
not exist.  rbgs is pinned to its canonical red-black path by running with
OMP threads > height/100 (pressure_solver.cpp:65)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import bind  # noqa: E402
from tests import cases  # noqa: E402


def norm(a):
    return float(np.sqrt((a.astype(np.float64) ** 2).sum()))


def main():
    bind.build(ref=True)
    R = bind.Ref()

    # 1. mgtest known-answer vector (mgtest.cpp:10-53 with .fill())
    u, rhs, flag, h, ref = cases.mgtest_problem(1025)
    R.canonical_threads(1025)
    mg = R.MG(1025, 1025)
    mg.set(u, rhs, flag)
    hist = [mg.residual(h)]
    for _ in range(5):
        mg.solve(h, False)
        hist.append(mg.residual(h))
    json.dump({"N": 1025, "residual_history": hist,
               "scaled_error": cases.mgtest_error(ref, mg.get_p()),
               "source": "oracle/_ref, OMP threads 11 (canonical rbgs path)"},
              open(os.path.join(HERE, "mgtest_kat.json"), "w"), indent=1)

    # 2. game level (ubootgl_app.hpp:29-30): flag mask, pyramid, 3 steps of norms
    flag = R.terrain_flag("/root/reference/resources/level2_hires2.png", 1)
    H, W = flag.shape
    R.canonical_threads(H)
    sim = R.Sim(flag, 0.8, 0.001)
    pyr = [sim.mg_flagc(l) for l in range(sim.mg_levels())]
    norms = []
    for _ in range(3):
        sim.step(0.001)
        norms.append({k: norm(sim.get(f)) for k, f in
                      (("vx", bind.VX), ("vy", bind.VY), ("p", bind.P), ("f", bind.F))})
    np.savez_compressed(
        os.path.join(HERE, "game_level.npz"),
        shape=np.array([H, W]),
        flag_bits=np.packbits(flag.astype(np.uint8)),
        **{f"pyr{l}_bits": np.packbits(a.astype(np.uint8)) for l, a in enumerate(pyr)},
        **{f"pyr{l}_shape": np.array(a.shape) for l, a in enumerate(pyr)},
        # coarse samples of the fields after step 3 (every 8th cell)
        vx_s=sim.get(bind.VX)[::8, ::8], vy_s=sim.get(bind.VY)[::8, ::8],
        p_s=sim.get(bind.P)[::8, ::8])
    json.dump({"W": W, "H": H, "dt": 0.001, "norms_after_step": norms,
               "levels": len(pyr)},
              open(os.path.join(HERE, "game_level_kat.json"), "w"), indent=1)

    # 3. stage-by-stage fields of one step on small grids (all advect quirks)
    for (W, H) in [(70, 40), (74, 44), (130, 97)]:
        c = cases.sim_case(W, H, seed=W * 100 + H)
        R.canonical_threads(H)
        s = R.Sim(c["flag"])
        s.set(bind.VX, c["vx"]); s.set(bind.VY, c["vy"])
        s.set(bind.VXB, c["vx"][::-1].copy()); s.set(bind.VYB, c["vy"][::-1].copy())
        s.set(bind.VX_ACCUM, c["vx_accum"]); s.set(bind.VY_ACCUM, c["vy_accum"])
        s.set(bind.P, c["p"])
        s.add_sink(0.4, 0.4 * H / W, 120.0)
        dt = float(s.dx)
        out = {}
        for name, st in (("accum", bind.ST_ACCUM), ("diffuse", bind.ST_DIFFUSE),
                         ("advect", bind.ST_ADVECT), ("setvbcs", bind.ST_SETVBCS),
                         ("project", bind.ST_PROJECT), ("setvbcs2", bind.ST_SETVBCS)):
            s.stage(st, dt)
            for k, f in (("vx", bind.VX), ("vy", bind.VY), ("vxb", bind.VXB), ("vyb", bind.VYB)):
                out[f"{name}_{k}"] = s.get(f)
        out["project_p"] = s.get(bind.P)
        out["project_f"] = s.get(bind.F)
        out["sinks_after"] = s.sinks()
        np.savez_compressed(os.path.join(HERE, f"stages_{W}x{H}.npz"), dt=np.float32(dt), **out)

    # 4. multigrid operators on an awkward size
    W, H = 67, 45
    flag, p, f = cases.random_fields(W, H, seed=42)
    rng = np.random.default_rng(43)
    flagc = (rng.random((H // 2, W // 2)) > 0.3).astype(np.float32)
    ec = rng.standard_normal((H // 2, W // 2)).astype(np.float32)
    R.canonical_threads(H)
    r, l2 = R.residual(p, f, flag, 0.01)
    m = R.MG(W, H); m.update_fields(flag); m.set(p, f, flag)
    m.solve(0.01, True)
    np.savez_compressed(os.path.join(HERE, "mg_ops_67x45.npz"),
                        rbgs3=R.rbgs(p, f, flag, 0.01, 1.0, 3), residual=r,
                        residual_l2=np.float32(l2), restrict=R.restrict(r),
                        prolongate=R.prolongate(ec, flagc, flag), vcycle_p=m.get_p(),
                        **{f"pyr{l}": m.flagc(l) for l in range(m.levels())})
    print("golden fixtures written to", HERE)
    os.system(f"ls -la {HERE}")


if __name__ == "__main__":
    main()
