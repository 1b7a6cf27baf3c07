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


FIELDS = (("vx", "VX", 0), ("vy", "VY", 1), ("p", "P", 0), ("vxc", "VX_CURRENT", 0), ("vyc", "VY_CURRENT", 1),
          ("vxb", "VXB", 0), ("f", "F", 0), ("ax", "VX_ACCUM", 0))


def check(W, H, steps, dt, rank, world, dev, verbose=True, skew=False):
    """Runs `steps` steps of the seeded W x H case as `world` row slabs (all ranks) and as ONE
    single-GPU Simulation (rank 0), and compares 8 fields bit for bit (up to the sign of zero)
    plus the residual norm.  Collective: every rank must call it.  Returns a dict (same on all
    ranks): bitwise_ok, fields, ranks, mismatches, exchanges, dist_levels, residual."""
    import ubootgl_b200 as u
    from ubootgl_b200 import capi, slab_boot

    c = cases.sim_case(W, H, seed=11, ndiscs=9, radius=H / 17.0)
    sinks = [[0.4, 0.4 * H / W, 120.0], [0.2, 0.7 * H / W, 60.0]]
    if skew:  # unequal slab heights (ubgl_slab_set_row_weights): rows get cheaper towards the top
        u.slab_set_row_weights(np.linspace(1.6, 0.6, H).astype(np.float32))
    try:
        plan = u.slab_plan(W, H, world, rank)
        S = u.SlabSimulation(c["flag"][plan["st_lo"]:plan["st_hi"]], W, H, rank, world,
                             slab_boot.blob_exchange(), device=dev)
    finally:
        if skew:
            u.slab_set_row_weights(None)
    for fld, key in ((capi.VX, "vx"), (capi.VY, "vy"), (capi.VX_ACCUM, "vx_accum"),
                     (capi.VY_ACCUM, "vy_accum"), (capi.P, "p")):
        S.set_from_global(fld, c[key])
    S.set_sinks(sinks)
    for _ in range(steps):
        S.step(dt)
    S.sync()
    got = {}
    for name, fid, short in FIELDS:
        lo, rows = S.get_own(getattr(capi, fid))
        got[name] = slab_boot.gather_rows(lo, rows, H - short)
    ssq = slab_boot.allreduce_sum(S.residual_sumsq())
    ex, hb = S.stats()
    slab_boot.barrier()  # every rank is done with its neighbours' arenas
    S.close()
    bad_fields = []
    resn = float(np.sqrt(ssq))
    res1 = resn
    if rank == 0:
        G = u.Simulation(c["flag"], device=dev)
        for fld, key in ((capi.VX, "vx"), (capi.VY, "vy"), (capi.VX_ACCUM, "vx_accum"),
                         (capi.VY_ACCUM, "vy_accum"), (capi.P, "p")):
            G.set(fld, c[key])
        G.set_sinks(sinks)
        for _ in range(steps):
            G.step(dt)
        for name, fid, short in FIELDS:
            a, b = got[name], G.get(getattr(capi, fid))
            same = ((a.view(np.uint32) == b.view(np.uint32)) | ((a == 0) & (b == 0)))
            if not same.all():
                bad_fields.append(name)
                bad = np.argwhere(~same)
                if verbose:
                    print(f"MISMATCH {name}: {len(bad)} cells, first {bad[:5].tolist()}, "
                          f"max abs {np.abs(a - b).max():.3e}, rel-L2 {cases.rel_l2(a, b):.3e}", flush=True)
        res1 = G.residual()
        G.close()
        if abs(resn - res1) > 1e-5 * max(res1, 1e-30):
            bad_fields.append("residual_norm")
            if verbose:
                print(f"MISMATCH residual norm: slabs {resn} single {res1}", flush=True)
    nbad = int(slab_boot.allreduce_max(float(len(bad_fields))))
    return {"bitwise_ok": nbad == 0, "fields": len(FIELDS), "ranks": world, "grid": [W, H], "steps": steps,
            "dt": dt, "rows": plan["own_hi"] - plan["own_lo"], "mismatched_fields": bad_fields if rank == 0 else None, "exchanges": ex,
            "halo_mb": hb / 1e6, "dist_levels": plan["dist_levels"], "residual": resn}


def main():
    import torch
    from ubootgl_b200 import slab_boot

    W = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    # dt = 0.02 is CFL ~ 25-75: back-traces leave the 16 ghost rows and are served by peer loads
    dt = float(sys.argv[4]) if len(sys.argv) > 4 else 0.002
    rank, world = slab_boot.init_distributed("gloo")
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    r = check(W, H, steps, dt, rank, world, dev, skew=len(sys.argv) > 5 and sys.argv[5] == "skew")
    if rank == 0:
        print(f"MGPU_EQUIV {'OK' if r['bitwise_ok'] else 'FAIL'} {W}x{H} ranks={world} steps={steps} dt={dt} "
              f"dist_levels={r['dist_levels']} exchanges={r['exchanges']} halo_MB={r['halo_mb']:.1f} "
              f"residual={r['residual']:.6g}", flush=True)
    slab_boot.shutdown()
    sys.exit(0 if r["bitwise_ok"] else 1)


if __name__ == "__main__":
    main()
