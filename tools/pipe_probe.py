"""Timeline probe of ubgl_sim_step_host_pipelined at 8192^2 (UBGL_PIPE_DEBUG=1 prints per-call phase times)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ubootgl_b200 as u
from ubootgl_b200 import capi
import bench
W, H, flag, vx, vy, dt = bench.make_inputs("channel8192")
sim = u.Simulation(flag, 0.8, 0.001)
sim.set(capi.VX, vx)
pin = lambda shape: torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
ax, ay = pin((H, W - 1)), pin((H - 1, W)); ax[:] = 0; ay[:] = 0
outs = dict(vx=pin((H, W - 1)), vy=pin((H - 1, W)), p=pin((H, W)), vx_current=pin((H, W - 1)), vy_current=pin((H - 1, W)))
for mode in ("full", "no_current", "no_accum", "p_only"):
    o = dict(outs)
    kw = dict(vx_accum=ax, vy_accum=ay)
    if mode in ("no_current", "p_only"): o.pop("vx_current"); o.pop("vy_current")
    if mode == "no_accum": kw = {}
    if mode == "p_only": o.pop("vx"); o.pop("vy")
    for _ in range(2): sim.step_host_pipelined(dt, **kw, **o)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); sim.step_host_pipelined(dt, **kw, **o); ts.append((time.perf_counter() - t0) * 1e3)
    sim.step_host_flush(**o)
    print(mode, "pipelined ms/call", [round(t, 1) for t in ts], flush=True)
