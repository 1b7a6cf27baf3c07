"""torchrun debugging aid: one slab step vs one single-GPU step, mismatch bounding boxes per field.
   tools/slab_diff.py W H [steps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import cases
import torch
import ubootgl_b200 as u
from ubootgl_b200 import capi, slab_boot
W = int(sys.argv[1]); H = int(sys.argv[2]); steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rank, world = slab_boot.init_distributed("gloo")
dev = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(dev)
plan = u.slab_plan(W, H, world, rank)
dt = float(np.float32(0.8) / np.float32(W - 1))
flag = cases.channel_flag_rows(W, H, plan["st_lo"], plan["st_hi"], seed=1234)
sim = u.SlabSimulation(flag, W, H, rank, world, slab_boot.blob_exchange(), device=dev)
vx = (flag[:, :-1] * flag[:, 1:]).astype(np.float32); vx[:, 0] = 1.0
sim.set(capi.VX, vx)
del vx
if rank == 0:
    print("plan", plan, flush=True)
G = None
if rank == 0:
    fl, _ = cases.channel_flag(W, H, seed=1234)
    G = u.Simulation(fl, device=dev)
    gx, gy = cases.uniform_stream(fl)
    G.set(capi.VX, gx); G.set(capi.VY, gy)
    del gx, gy
for s in range(steps):
    sim.step(dt); sim.sync()
    if G is not None:
        G.step(dt)
    for name, fld in (("vxb(advected pre-proj? no: back)", capi.VXB), ("f", capi.F), ("p", capi.P), ("vx", capi.VX), ("vy", capi.VY)):
        lo, a = sim.get_own(fld)
        full = slab_boot.gather_rows(lo, a, H - 1 if fld == capi.VY else H)
        if rank == 0:
            b = G.get(fld)
            bad = ~((full.view(np.uint32) == b.view(np.uint32)) | ((full == 0) & (b == 0)))
            if bad.any():
                ys, xs = np.nonzero(bad)
                print(f"step {s} {name}: {bad.sum()} bad cells, rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()} "
                      f"maxabs {np.nanmax(np.abs(full - b)):.3g}; first rows {np.unique(ys)[:12].tolist()} "
                      f"first cols {np.unique(xs)[:12].tolist()}", flush=True)
                print("   cells (y,x,slab,single):", [(int(y), int(x), float(full[y, x]), float(b[y, x])) for y, x in zip(ys[:30], xs[:30])], flush=True)
            else:
                print(f"step {s} {name}: identical", flush=True)
