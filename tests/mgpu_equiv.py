"""Launched by torchrun (one process per GPU): the row-slab decomposed step must
reproduce the single-GPU step BIT FOR BIT (same kernels, same per-cell
arithmetic; only the tile origins and the halo plumbing differ).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/mgpu_equiv.py [W H steps dt]
Rank 0 prints "MGPU_EQUIV OK ..." and exits 0 on success.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import cases  # noqa: E402


def main():
    import torch
    import ubootgl_b200 as u
    from ubootgl_b200 import capi, slab_boot

    W = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    rank, world = slab_boot.init_distributed("gloo")
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)

    c = cases.sim_case(W, H, seed=11, ndiscs=9, radius=H / 17.0)
    # dt = 0.02 is CFL ~ 25-75: back-traces leave the 16 ghost rows and are served by peer loads
    dt = float(sys.argv[4]) if len(sys.argv) > 4 else 0.002
    sinks = [[0.4, 0.4 * H / W, 120.0], [0.2, 0.7 * H / W, 60.0]]
    plan = u.slab_plan(W, H, world, rank)
    S = u.SlabSimulation(c["flag"][plan["st_lo"]:plan["st_hi"]], W, H, rank, world,
                         slab_boot.blob_exchange(), device=dev)
    for fld, key in ((capi.VX, "vx"), (capi.VY, "vy"), (capi.VX_ACCUM, "vx_accum"),
                     (capi.VY_ACCUM, "vy_accum"), (capi.P, "p")):
        S.set_from_global(fld, c[key])
    S.set_sinks(sinks)
    for _ in range(steps):
        S.step(dt)
    S.sync()
    got = {}
    for name, fld, hh in (("vx", capi.VX, H), ("vy", capi.VY, H - 1), ("p", capi.P, H),
                          ("vxc", capi.VX_CURRENT, H), ("vyc", capi.VY_CURRENT, H - 1),
                          ("vxb", capi.VXB, H), ("f", capi.F, H), ("ax", capi.VX_ACCUM, H)):
        lo, rows = S.get_own(fld)
        got[name] = slab_boot.gather_rows(lo, rows, hh)
    ssq = slab_boot.allreduce_sum(S.residual_sumsq())
    ex, hb = S.stats()
    ok = True
    if rank == 0:
        G = u.Simulation(c["flag"], device=dev)
        for fld, key in ((capi.VX, "vx"), (capi.VY, "vy"), (capi.VX_ACCUM, "vx_accum"),
                         (capi.VY_ACCUM, "vy_accum"), (capi.P, "p")):
            G.set(fld, c[key])
        G.set_sinks(sinks)
        for _ in range(steps):
            G.step(dt)
        ref = {"vx": G.get(capi.VX), "vy": G.get(capi.VY), "p": G.get(capi.P),
               "vxc": G.get(capi.VX_CURRENT), "vyc": G.get(capi.VY_CURRENT), "vxb": G.get(capi.VXB),
               "f": G.get(capi.F), "ax": G.get(capi.VX_ACCUM)}
        res1 = G.residual()
        for k in ref:
            a, b = got[k], ref[k]
            same = ((a.view(np.uint32) == b.view(np.uint32)) | ((a == 0) & (b == 0)))
            if not same.all():
                ok = False
                bad = np.argwhere(~same)
                print(f"MISMATCH {k}: {len(bad)} cells, first {bad[:5].tolist()}, "
                      f"max abs {np.abs(a - b).max():.3e}, rel-L2 {cases.rel_l2(a, b):.3e}", flush=True)
        resn = float(np.sqrt(ssq))
        if abs(resn - res1) > 1e-5 * max(res1, 1e-30):
            ok = False
            print(f"MISMATCH residual norm: slabs {resn} single {res1}", flush=True)
        print(f"MGPU_EQUIV {'OK' if ok else 'FAIL'} {W}x{H} ranks={world} steps={steps} dt={dt} "
              f"dist_levels={plan['dist_levels']} exchanges={ex} halo_MB={hb / 1e6:.1f} "
              f"residual={resn:.6g}", flush=True)
    okt = slab_boot.allreduce_max(0.0 if ok else 1.0)
    sys.exit(0 if okt == 0.0 else 1)


if __name__ == "__main__":
    main()
