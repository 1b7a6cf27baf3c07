import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import cases
import torch
import ubootgl_b200 as u
from ubootgl_b200 import capi, slab_boot
import bench
S = int(sys.argv[1]); steps = int(sys.argv[2])
rank, world = slab_boot.init_distributed("gloo")
dev = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(dev)
W = H = S
plan = u.slab_plan(W, H, world, rank)
dt = float(np.float32(0.8) / np.float32(W - 1))
flag = cases.channel_flag_rows(W, H, plan["st_lo"], plan["st_hi"], seed=1234)
sim = u.SlabSimulation(flag, W, H, rank, world, slab_boot.blob_exchange(), device=dev)
vx = (flag[:, :-1] * flag[:, 1:]).astype(np.float32); vx[:, 0] = 1.0
sim.set(capi.VX, vx)
G = None
if rank == 0 and S <= 16384:
    fl, _ = cases.channel_flag(W, H, seed=1234)
    G = u.Simulation(fl, device=dev)
    gx, gy = cases.uniform_stream(fl)
    G.set(capi.VX, gx); G.set(capi.VY, gy)
for s in range(steps):
    try:
        sim.step(dt); sim.sync()
        err = ""
    except Exception as e:
        err = str(e)
    lo, a = sim.get_own(capi.VX); _, b = sim.get_own(capi.VY); _, pp = sim.get_own(capi.P)
    msg = f"step {s} rank {rank}: max|vx| {np.nanmax(np.abs(a)):.4g} max|vy| {np.nanmax(np.abs(b)):.4g} max|p| {np.nanmax(np.abs(pp)):.4g} nan {int(np.isnan(a).sum())} {err}"
    if G is not None:
        G.step(dt)
        ga = G.get(capi.VX)[plan['own_lo']:plan['own_hi']]; gb = G.get(capi.VY)[plan['own_lo']:plan['own_hi']]
        msg += f" | single: max|vx| {np.abs(ga).max():.4g} max|vy| {np.abs(gb).max():.4g} diffvx {np.abs(ga-a).max():.3g}"
    print(msg, flush=True)
