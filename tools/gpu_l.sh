#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
for T in 8 16 32; do
echo "== host threads $T"
UBGL_HOST_THREADS=$T timeout 300 python tools/pipe_probe.py 2>&1 | grep -v "^\[pipe\]" | head -2
UBGL_HOST_THREADS=$T timeout 300 python - <<PY
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import ubootgl_b200 as u
from ubootgl_b200 import capi
import bench
W, H, flag, vx, vy, dt = bench.make_inputs("channel8192")
sim = u.Simulation(flag, 0.8, 0.001); sim.set(capi.VX, vx)
pin = lambda shape: torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
ax, ay = pin((H, W - 1)), pin((H - 1, W)); ax[:] = 0; ay[:] = 0
outs = dict(vx=pin((H, W - 1)), vy=pin((H - 1, W)), p=pin((H, W)), vx_current=pin((H, W - 1)), vy_current=pin((H - 1, W)))
sim.step_host(dt, vx_accum=ax, vy_accum=ay, **outs)
ts=[]
for _ in range(6):
    t0=time.perf_counter(); sim.step_host(dt, vx_accum=ax, vy_accum=ay, **outs); ts.append((time.perf_counter()-t0)*1e3)
print("sync step_host ms", [round(t,1) for t in ts])
PY
done
