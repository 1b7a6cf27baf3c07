"""Profiling harness for ncu: one warm-up step, then `--steps` Simulation::step
between cudaProfilerStart/Stop (run ncu with --profile-from-start off).
    ncu --profile-from-start off ... python tools/prof_step.py --size 8192 --steps 1
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ctypes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="channel8192")
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
a = ap.parse_args()

import ubootgl_b200 as u  # noqa: E402
from ubootgl_b200 import capi  # noqa: E402

W, H, flag, vx, vy, dt = bench.make_inputs(a.workload)
sim = u.Simulation(flag, bench.PWIDTH, bench.MU)
sim.set(capi.VX, vx)
sim.set(capi.VY, vy)
for _ in range(a.warmup):
    sim.step(dt)
sim.sync()
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaProfilerStart()
for _ in range(a.steps):
    sim.step(dt)
sim.sync()
rt.cudaProfilerStop()
print("profiled", a.steps, "steps of", a.workload)
