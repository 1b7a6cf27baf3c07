#!/usr/bin/env python
"""bench.py -- fluid-step throughput (MLUP/s) of the B200-native hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

A "step" is one Simulation::step (accumulate, diffuse, advect, divergence,
2 multigrid V-cycles, gradient, save) on synthetic input.  Workloads
(SURVEY.md section 8d / BASELINE.json configs):

  channel8192  8192x8192 channel with 32 obstacles, uniform stream, dt = h
               (configs[2]; the config BASELINE.json's target is quoted on) -- default at N=1
  channel32768 32768x32768, same generator, row-slab decomposed (configs[3]) -- default at N>1
  game         1090x436-like game level (configs[1]) -- L2 resident, latency bound
  explosion4096  configs[4]: 4096^2 channel; every step 16 craters (Terrain::drawCircle, diam 24)
               carved on the device + coarse-flag rebuild + their pressure sinks, the fluid
               step, 1 M fluid tracers and 1 M floating items
  channel<S>   any other square size S

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import cases  # noqa: E402  (seeded synthetic generators, numpy only)

VCYCLES = 2                      # simulation.cpp:189-190
B_NONMG = 152.0                  # algorithmic B/cell of the non-MG stages
B_VCYCLE = 186.7                 # algorithmic B/level-0-cell per V-cycle, whole pyramid
PWIDTH, MU = 0.8, 0.001


def pcie_d2h_bytes(outs):
    """Bytes ubgl_sim_step_host moves device -> host: fields of 16 MB and more send vx / vy once and
    fill the *_current mirrors by host copies; smaller ones are simply downloaded too."""
    return sum(a.nbytes for k_, a in outs.items() if not (k_.endswith("_current") and a.nbytes >= (16 << 20)))


def bytes_per_cell(k=VCYCLES):
    return B_NONMG + B_VCYCLE * k


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_dims(name):
    if name == "game":
        return 1090, 436
    if name.startswith("explosion"):
        s = int(name[len("explosion"):])
        return s, s
    if name.startswith("channel"):
        s = int(name[len("channel"):])
        return s, s
    raise SystemExit(f"unknown workload {name}")


def make_inputs(name):
    W, H = workload_dims(name)
    if name == "game":
        from tests import golden_util
        flag, _, _ = golden_util.game_level()
        vx = np.zeros((H, W - 1), np.float32)
        vx[:, 0] = 1.0
        vy = np.zeros((H - 1, W), np.float32)
        dt = 0.001
    else:
        flag, _ = cases.channel_flag(W, H, seed=1234)  # explosion<S> shares the channel generator
        vx, vy = cases.uniform_stream(flag)
        dt = float(np.float32(PWIDTH) / np.float32(W - 1))  # dt = h, CFL ~ 1
    return W, H, flag, vx, vy, dt


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  Sampled in-process through
    NVML (what nvidia-smi itself reads; spawning nvidia-smi every 100 ms perturbs frame loops
    that synchronise with the device), falling back to the nvidia-smi query of the recipe."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
    BITS = (0x8, 0x40, 0x20, 0x4)  # nvmlClocksEventReason{HwSlowdown,HwThermalSlowdown,SwThermalSlowdown,SwPowerCap}

    def __init__(self, index=0):
        self.rows = []   # (sm_mhz, sm_max_mhz, set of reasons)
        self.stop = False
        self.index = index
        self.nvml = None
        self.source = "nvidia-smi"
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml, self.h, self.source = pynvml, h, "nvml"
        except Exception:
            self.nvml = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self.h))
        self.rows.append((sm, self.max_mhz, {nm for nm, b in zip(self.NAMES, self.BITS) if mask & b}))

    def sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
        r = [c.strip() for c in out.strip().split(",")]
        if len(r) >= 7:
            self.rows.append((float(r[0]), float(r[1]),
                              {nm for nm, v in zip(self.NAMES, r[3:7]) if v.lower().startswith("active")}))

    def run(self):
        while not self.stop:
            try:
                self.sample_nvml() if self.nvml else self.sample_smi()
            except Exception:
                pass
            time.sleep(0.05 if self.nvml else 0.1)

    def sample_once(self):
        """One sample from the calling thread (the multi-GPU runs: see slab_bench.py)."""
        try:
            self.sample_nvml() if self.nvml else self.sample_smi()
        except Exception:
            pass

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        reasons = set()
        for r in self.rows:
            reasons |= r[2]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


# ---------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own OpenMP+AVX2 CPU solver
# ---------------------------------------------------------------------------
def cpu_reference_run(name, steps, warmup, threads=None):
    """Times Simulation::step of the unmodified reference (oracle/_ref when it
    was built from /root/reference, else the C restatement) on the host cores."""
    from oracle import bind as ob
    if ob.have_ref():
        chk, kind = ob.Ref(), "reference"
    else:
        if not ob.have_port():
            ob.build(ref=False)
        chk, kind = ob.Port(), "port"
    nproc = chk.num_procs()
    threads = threads or nproc
    chk.set_threads(threads)
    W, H, flag, vx, vy, dt = make_inputs(name)
    sim = chk.Sim(flag, PWIDTH, MU)
    sim.set(ob.VX, vx)
    sim.set(ob.VY, vy)
    for _ in range(warmup):
        sim.step(dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.step(dt)
    t = time.perf_counter() - t0
    path = "pipelined path 3 (H/T >= 100)" if H // threads >= 100 else "canonical red-black"
    return dict(value=W * H * steps / t / 1e6, ms_per_step=t / steps * 1e3, cores=threads, kind=kind,
                nproc=nproc, rbgs_path=path, W=W, H=H, steps=steps, warmup=warmup)


def cpu_reference_isolated(name, steps, warmup, threads=None):
    """cpu_reference_run in a child process: the reference's pipelined rbgs path 3 reads one row
    past the end of its grids, which can end a process depending on the heap
    layout; the bench line must survive that.  Returns the dict, or None if the child died."""
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", name, str(steps), str(warmup),
           str(threads or 0)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        for line in reversed(out.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
    except Exception:
        pass
    return None


def l2_note(N):
    return (f"state {12 * N * 4 / 1e6:.0f} MB per step vs 126 MB L2"
            + (" (inputs larger than L2)" if N * 4 > 126e6 else " (L2 RESIDENT: latency bound, HBM % not meaningful)"))


def workload_config(ran, asked=None, parallelism="1 GPU"):
    """`config` of a bench line: the SAME keys (and, for the same workload, the same values) in
    our arm and in the reference arm.  `ran` is the workload that was timed; `sampled_for` names
    the workload it stands in for when the asked one does not fit the arm (the reference CPU run
    of channel32768 needs ~80 GB of host memory and minutes per step)."""
    W, H = workload_dims(ran)
    dt = 0.001 if ran == "game" else float(np.float32(PWIDTH) / np.float32(W - 1))
    return {"workload": ran, "grid": [W, H], "vcycles_per_step": VCYCLES, "dt": dt,
            "sampled_for": asked if (asked and asked != ran) else None,
            "parallelism": parallelism, "l2": l2_note(W * H)}


def cpu_baseline_entry(sample, steps=2, warmup=1, policy=True):
    """cpu_baseline object: the reference's own solver on this box's host cores, bounded sample."""
    r = cpu_reference_isolated(sample, steps, warmup) or cpu_reference_isolated(sample, steps, warmup)  # one retry
    if r is None:
        return {"value": None, "unit": "MLUP/s", "cores": None, "kind": "reference",
                "sample": f"{sample}: the reference process died twice (its rbgs path 3 reads out of bounds)"}
    cpu = {"value": r["value"], "unit": "MLUP/s", "cores": r["cores"], "kind": r["kind"],
           "sample": f"{sample}: {r['warmup']} warm-up + {r['steps']} timed Simulation::step "
                     f"({r['ms_per_step']:.0f} ms/step), OMP threads {r['cores']} of {r['nproc']}, rbgs {r['rbgs_path']}"}
    # the game's own thread policy, sim_loop.cpp:19: max(1, nproc/2 - 1)
    tg = max(1, r["nproc"] // 2 - 1)
    if policy and tg != r["cores"]:
        r2 = cpu_reference_isolated(sample, 1 if sample != "game" else 20, 1, threads=tg)
        if r2:
            cpu["game_thread_policy"] = {"value": r2["value"], "cores": tg, "ms_per_step": r2["ms_per_step"],
                                         "rbgs_path": r2["rbgs_path"]}
    return cpu


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    asked = args.workload or ("channel8192" if args.gpus == 1 else "channel32768")
    sample = asked
    W, H = workload_dims(asked)
    if W * H > 8192 * 8192:
        # a 32768^2 reference step needs ~80 GB and minutes per step: time the
        # 8192^2 member of the same generator instead and SAY so (config.workload = what ran)
        sample = "channel8192"
    W, H = workload_dims(sample)
    # exactly the K timed steps after W warm-up steps the driver asked for, as long as that stays
    # a bounded sample (~0.55 s per 8192^2 step on 16 cores); what actually ran is what is printed
    est = 0.6 * (W * H) / (8192.0 * 8192.0)
    steps = max(1, min(args.steps, int(120.0 / max(est, 1e-3))))
    warmup = max(1, min(args.warmup, 5))
    r = cpu_reference_isolated(sample, steps, warmup) or cpu_reference_isolated(sample, steps, warmup)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "the reference process died twice (out-of-bounds "
                          "read in its pipelined rbgs path, SURVEY.md A.3)"}), flush=True)
        return
    line = {
        "impl": "reference", "metric": "fluid_step_throughput", "value": r["value"], "unit": "MLUP/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(sample, asked, parallelism="1 GPU" if args.gpus == 1 else f"row slabs x{args.gpus}"),
        "note": f"reference CPU solver (OpenMP+AVX2, oracle/_ref = the unmodified reference) on the GPU box's host "
                f"cores: {r['cores']} OMP threads; the same CPU run stands beside every GPU count",
        "cpu_baseline": {"value": r["value"], "unit": "MLUP/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": f"{sample}: {r['warmup']} warm-up + {r['steps']} timed Simulation::step, "
                                   f"OMP threads {r['cores']} of {r['nproc']} procs, rbgs {r['rbgs_path']}"},
        "e2e": {"value": r["value"], "unit": "MLUP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def run_single_gpu(args, name):
    import torch
    import ubootgl_b200 as u
    from ubootgl_b200 import capi

    if u.lib.ubgl_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible to libubgl.so (there is no CPU fallback)")
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    W, H, flag, vx, vy, dt = make_inputs(name)
    N = W * H
    sim = u.Simulation(flag, PWIDTH, MU, device=dev)
    if os.environ.get("UBGL_BENCH_GRAPH") == "0":  # A/B of UBGL_OPT_GRAPH (small grids only)
        sim.set_option(capi.OPT_GRAPH, 0)
    sim.set(capi.VX, vx)
    sim.set(capi.VY, vy)
    stream = torch.cuda.ExternalStream(sim.stream(), device=dev)
    K, Wm = args.steps, args.warmup

    # ---- device-resident throughput (`value`) ----
    for _ in range(Wm):
        sim.step(dt)
    sim.sync()
    torch.cuda.synchronize()
    l0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev) as clk:
        e0.record(stream)
        for _ in range(K):
            sim.step(dt)
        e1.record(stream)
        sim.sync()
        torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    launches = sim.launch_count() - l0
    ms_step = ms_total / K
    value = N * K / (ms_total * 1e-3) / 1e6
    res_after = sim.residual()

    # ---- MG V-cycle latency (BASELINE.json metric: "MG V-cycle ms"), resident p/f/flag ----
    KV = 10  # mgtest.cpp:55-61 times 10 solves
    sim.mg_solve(2)
    sim.sync()
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    v0.record(stream)
    sim.mg_solve(KV)
    v1.record(stream)
    sim.sync()
    torch.cuda.synchronize()
    vc_ms = v0.elapsed_time(v1) / KV

    # ---- per-kernel profile (CUDA events around every launch, same stream) ----
    sim.profile(True)
    PK = min(K, 5)
    for _ in range(PK):
        sim.step(dt)
    sim.sync()
    stats = sim.kernel_stats()
    sim.profile(False)
    prof_total = sum(ms for _, ms in stats.values())
    kern = sorted(((ms / PK, n // PK, k, l) for (k, l), (n, ms) in stats.items()), reverse=True)
    peak, peak_src = peaks()
    roof = dominant_roofline(kern, W, H, peak, peak_src, prof_total / PK, workload=name)

    # ---- end to end through the host-mirror API (`e2e`) ----
    pin = lambda shape: torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
    ax, ay = pin((H, W - 1)), pin((H - 1, W))
    ax[:] = 0
    ay[:] = 0
    outs = dict(vx=pin((H, W - 1)), vy=pin((H - 1, W)), p=pin((H, W)),
                vx_current=pin((H, W - 1)), vy_current=pin((H - 1, W)))
    h2d = ax.nbytes + ay.nbytes
    d2h = pcie_d2h_bytes(outs)
    KE = max(2, min(K, 8))
    sim.step_host(dt, vx_accum=ax, vy_accum=ay, **outs)
    t_calls = []
    for _ in range(KE):
        t0 = time.perf_counter()
        sim.step_host(dt, vx_accum=ax, vy_accum=ay, **outs)
        t_calls.append(time.perf_counter() - t0)
    t_e2e = sum(t_calls) / KE   # mean of KE synchronous calls (each returns with the mirrors filled)
    e2e = N / t_e2e / 1e6

    # ---- the optional pipelined form of the same call (mirrors one step late): extra key, not the headline ----
    pipe = None
    try:
        sim.step_host_pipelined(dt, vx_accum=ax, vy_accum=ay, **outs)  # fills the pipeline
        sim.step_host_pipelined(dt, vx_accum=ax, vy_accum=ay, **outs)
        t_p = []
        for _ in range(KE):
            t0 = time.perf_counter()
            sim.step_host_pipelined(dt, vx_accum=ax, vy_accum=ay, **outs)
            t_p.append(time.perf_counter() - t0)
        sim.step_host_flush(**outs)
        pipe = {"value": N / (sum(t_p) / KE) / 1e6, "unit": "MLUP/s", "ms_per_step": sum(t_p) / KE * 1e3,
                "ms_min": min(t_p) * 1e3, "ms_max": max(t_p) * 1e3, "calls": KE,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "ubgl_sim_step_host_pipelined: same bytes over PCIe per step, the download of step n-1 overlaps "
                       "the upload and the compute of step n; the mirrors a call returns are one step late "
                       "(the reference's render thread reads them that way)"}
    except Exception as e:
        pipe = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- CPU baseline: the reference's own solver on this box's host cores ----
    cpu = None
    if not args.no_cpu_baseline:
        sample = name if N <= 8192 * 8192 else "channel8192"
        cpu = cpu_baseline_entry(sample, steps=3 if sample != "game" else 20, warmup=1)

    # ---- the 1-GPU base of the strong-scaling study (BASELINE.json: ">= 6x at 8 GPUs on 32768^2"):
    # the SAME workload the N > 1 runs decompose, whole on this one GPU with the single-GPU kernels,
    # measured in this process so that the driver's N = 1 record contains it ----
    base = None
    if name == "channel8192" and not args.no_strong_base:
        del sim, outs, ax, ay
        base = strong_scaling_base("channel32768", dev, steps=min(K, 5), warmup=3)

    bpc = bytes_per_cell()
    step_gbs = N * bpc / (ms_step * 1e-3) / 1e9
    line = {
        "metric": "fluid_step_throughput", "value": value, "unit": "MLUP/s", "n_gpus": 1, "steps": K,
        "warmup": Wm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name),
        "run_info": {"residual_after": res_after, "fused": True,
                     "scaling_note": "N = 1 anchors the strong-scaling study of channel32768 (see strong_scaling_base); "
                                     "MLUP/s is size-normalised, so the driver's v_N / (N v_1) compares like with like"},
        "strong_scaling_base": base,
        "roofline": roof,
        "roofline_by_kernel": roofline_by_kernel(kern, W, H, peak, prof_total / PK, workload=name),
        "roofline_step": {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s",
                          "frac": step_gbs / peak, "bytes_per_cell": bpc,
                          "model": "stage-wise algorithmic bytes 152 + 186.7*k B/cell (SURVEY.md 8d), k=2",
                          "peak_source": peak_src},
        "vcycle": {"ms": vc_ms, "algorithmic_gbs": N * B_VCYCLE / (vc_ms * 1e-3) / 1e9,
                   "frac": N * B_VCYCLE / (vc_ms * 1e-3) / 1e9 / peak,
                   "note": "one MG::solve V(3,3) on the resident level-0 fields, mean of 10 (mgtest.cpp:55-61); "
                           "186.7 algorithmic B per level-0 cell"},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e, "unit": "MLUP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": t_e2e * 1e3, "calls": KE, "ms_min": min(t_calls) * 1e3, "ms_max": max(t_calls) * 1e3,
                "api": "ubgl_sim_step_host (pinned host mirrors: accumulators in; vx, vy, p out over PCIe; "
                       + ("vx_current, vy_current filled from the vx, vy mirrors by host threads like "
                          "saveCurrentVelocityFields' memcpy)" if N * 4 >= (16 << 20) else
                          "vx_current, vy_current downloaded too: fields under 16 MB)")},
        "e2e_pipelined": pipe,
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "kernels_ms_per_step": [{"kernel": k, "level": l, "launches": n, "ms": round(ms, 4)} for ms, n, k, l in kern[:12]],
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# configs[4]: explosion-heavy step -- terrain edits + step + tracers + items
# ---------------------------------------------------------------------------
N_PARTICLES = 1_000_000
GAME_DT = 1.0 / 60.0          # one rendered frame (items sub-step up to 15x, advect_floating_items.cpp:183)
CRATERS, CRATER_DIAM = 16, 24


def explosion_inputs(name, n_particles):
    from tests import next_cases
    W, H, flag, vx, vy, dt = make_inputs(name)
    items = next_cases.make_items(n_particles, W, H, seed=7, flag=None, cluster=0.0)
    rng = np.random.default_rng(7)
    st = next_cases.tracer_state(n_particles, 30)
    ys, xs = np.nonzero(flag[2:-2, 2:-2])
    pick = rng.integers(0, len(ys), n_particles)
    cell = np.float32(PWIDTH) / np.float32(W)
    st["points"][:, 0, 0] = (xs[pick] + 2.5) * cell      # uniformly in fluid cells
    st["points"][:, 0, 1] = (ys[pick] + 2.5) * cell
    st["ages"][:] = (rng.random(n_particles) * 6.0).astype(np.float32)
    return W, H, flag, vx, vy, dt, items, st


def crater_list(g, W, H):
    out = []
    for _ in range(CRATERS):
        cx = CRATER_DIAM + 2 + g.u() * (W - 2 * CRATER_DIAM - 5)
        cy = CRATER_DIAM + 2 + g.u() * (H - 2 * CRATER_DIAM - 5)
        out.append((cx, cy, CRATER_DIAM))
    return np.asarray(out, np.float32)


def run_explosion(args, name):
    import torch
    import ubootgl_b200 as u
    from ubootgl_b200 import capi

    if u.lib.ubgl_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible to libubgl.so (there is no CPU fallback)")
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    W, H, flag, vx, vy, dt, items, st = explosion_inputs(name, N_PARTICLES)
    N = W * H
    h = float(np.float32(PWIDTH) / np.float32(W - 1))
    sim = u.Simulation(flag, PWIDTH, MU, device=dev)
    sim.set(capi.VX, vx)
    sim.set(capi.VY, vy)
    T = u.Tracers(N_PARTICLES, 30, device=dev)
    T.set_state(st)
    I = u.Items(items, device=dev)
    stream = torch.cuda.ExternalStream(sim.stream(), device=dev)
    g = cases.LCG(99)
    K, Wm = args.steps, args.warmup
    seed = [12345]

    def frame(host_out=None):
        circ = crater_list(g, W, H)
        sim.draw_circles(circ, 1.0)                       # explosion.cpp:60 + ubootgl_app.cpp:111-112
        new = np.concatenate([circ[:, :2] * np.float32(h), np.full((CRATERS, 1), 120.0, np.float32)], 1)
        sim.set_sinks(np.concatenate([sim.sinks(), new]))  # explosion.cpp:33 (one sink per crater here)
        if host_out is None:
            sim.step(dt)
        else:
            sim.step_host(dt, **host_out)
        seed[0] = (1103515245 * seed[0] + 12345) & 0x7FFFFFFF
        T.advect(sim, GAME_DT, seed[0])
        I.advect_simple(sim, GAME_DT)

    for _ in range(Wm):
        frame()
    sim.sync()
    torch.cuda.synchronize()
    l0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev) as clk:
        e0.record(stream)
        for _ in range(K):
            frame()
        e1.record(stream)
        sim.sync()
        torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / K
    launches = sim.launch_count() - l0

    sim.profile(True)
    PK = min(K, 5)
    for _ in range(PK):
        frame()
    sim.sync()
    stats = sim.kernel_stats()
    sim.profile(False)
    per = {}
    for (k, l), (n, ms) in stats.items():
        per[k] = per.get(k, 0.0) + ms / PK
    kern = sorted(((ms / PK, n // PK, k, l) for (k, l), (n, ms) in stats.items()), reverse=True)
    peak, peak_src = peaks()
    fluid = [r for r in kern if r[2] not in ("tracers", "items", "terrain", "colocate", "coarsen_flag")]
    roof = dominant_roofline(fluid, W, H, peak, peak_src, sum(per.values()))

    # end to end: crater + sink lists in, fields + item records out, every frame
    pin = lambda shape, dt_=torch.float32: torch.empty(shape, dtype=dt_, pin_memory=True).numpy()
    outs = dict(vx=pin((H, W - 1)), vy=pin((H - 1, W)), p=pin((H, W)),
                vx_current=pin((H, W - 1)), vy_current=pin((H - 1, W)))
    KE = max(2, min(K, 8))
    rec = torch.empty(N_PARTICLES * capi.ITEM_DTYPE.itemsize, dtype=torch.uint8,
                      pin_memory=True).numpy().view(capi.ITEM_DTYPE)   # pinned item mirror
    frame(outs); I.get(rec)
    t_calls = []
    for _ in range(KE):
        t0 = time.perf_counter()
        frame(outs)
        I.get(rec)
        t_calls.append(time.perf_counter() - t0)
    t_e2e = sum(t_calls) / KE
    d2h = pcie_d2h_bytes(outs) + rec.nbytes
    h2d = CRATERS * 12 + len(sim.sinks()) * 12

    # What bounds the item pass (the largest single item of this frame).  The reference tests an item against every
    # position of its x-bin (advect_floating_items.cpp:167-181: N^2 / 100 pair tests); the kernel walks the y-cells
    # y-1, y, y+1 of the x-bin only (next.cu).  Pair tests per frame are counted here on the host from the downloaded
    # records with the kernel's own binning; the ceiling is the SM issue rate at ~12 instructions per test.
    def items_roofline():
        pos, size = rec["pos"].astype(np.float64), rec["size"].astype(np.float64)
        r2max = float((size[:, 0] * size[:, 1] * 0.4).max())
        yrange = PWIDTH * H / W
        NY = 2048
        cy = max(np.sqrt(r2max) * 1.0001, yrange / NY)
        xb = (pos[:, 0] * 100.0).astype(np.int64) % 100
        yc = np.clip(np.floor(pos[:, 1] / cy), 0, NY - 1).astype(np.int64)
        occ = np.zeros((100, NY + 2), np.int64)
        np.add.at(occ, (xb, yc + 1), 1)
        near = occ[:, :-2] + occ[:, 1:-1] + occ[:, 2:]        # entries of cells y-1, y, y+1
        pairs = int(near[xb, yc].sum())
        ms = per.get("items", float("nan"))
        props = torch.cuda.get_device_properties(dev)
        issue = props.multi_processor_count * 4 * 32 * 1.965e9   # thread-instructions per second at 1965 MHz
        ceil_pairs = issue / 12.0
        rate = pairs / (ms * 1e-3)
        return {"bound": "latency (dependent walk of three sorted y-cells per item, a handful of contacts each); neither HBM "
                         "nor issue bound", "kernel": "items (r2max, keys, radix sort, offsets, k_items_advect)",
                "ms_per_frame": ms, "pair_tests_per_frame": pairs, "pair_tests_reference_scheme": int((occ.sum(axis=1) ** 2).sum()),
                "achieved": rate / 1e9, "peak": ceil_pairs / 1e9, "unit": "G pair tests/s", "frac": rate / ceil_pairs,
                "peak_model": "SM issue rate (SMs x 4 schedulers x 32 lanes x 1.965 GHz) / 12 instructions per pair test",
                "record_bytes_gbs": N_PARTICLES * 2 * capi.ITEM_DTYPE.itemsize / (ms * 1e-3) / 1e9,
                "record_bytes_frac_of_hbm": N_PARTICLES * 2 * capi.ITEM_DTYPE.itemsize / (ms * 1e-3) / 1e9 / peak}
    try:
        roof_items = items_roofline()
    except Exception as e:  # never lose the bench line over the diagnostic
        roof_items = {"error": repr(e)}

    cpu = None
    if not args.no_cpu_baseline:  # in a child process, like cpu_reference_isolated
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-worker-explosion", name],
                                 capture_output=True, text=True, timeout=900)
            cpu = next((json.loads(l) for l in reversed(out.stdout.strip().splitlines()) if l.startswith("{")), None)
        except Exception:
            cpu = None
        if cpu is None:
            cpu = {"value": None, "unit": "MLUP/s", "cores": None, "kind": "reference",
                   "sample": f"{name}: the reference process died (out-of-bounds read in its pipelined rbgs path)"}
    bpc = bytes_per_cell()
    step_only = sum(ms for ms, n, k, l in fluid)
    line = {
        "metric": "fluid_step_throughput", "value": N / (ms_step * 1e-3) / 1e6, "unit": "MLUP/s", "n_gpus": 1,
        "steps": K, "warmup": Wm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name),
        "run_info": {"per_step": f"{CRATERS} craters diam {CRATER_DIAM} + coarse-flag/mask rebuild, {CRATERS} sinks, "
                                 f"Simulation::step, {N_PARTICLES} tracers x30 ring, {N_PARTICLES} simple items "
                                 f"(game dt 1/60, up to 15 sub-steps, force scatter by atomicAdd)"},
        "particles": {"tracers_per_s": N_PARTICLES / (per.get("tracers", float("nan")) * 1e-3),
                      "items_per_s": N_PARTICLES / (per.get("items", float("nan")) * 1e-3),
                      "tracers_ms": per.get("tracers"), "items_ms": per.get("items"),
                      "terrain_ms": per.get("terrain", 0.0) + per.get("coarsen_flag", 0.0) + per.get("other", 0.0),
                      "fluid_step_ms": step_only,
                      "fluid_step_mlups": N / (step_only * 1e-3) / 1e6},
        "roofline": roof,
        "roofline_items": roof_items,
        "roofline_step": {"bound": "hbm", "achieved": N * bpc / (step_only * 1e-3) / 1e9, "peak": peak,
                          "unit": "GB/s", "frac": N * bpc / (step_only * 1e-3) / 1e9 / peak, "bytes_per_cell": bpc,
                          "model": "fluid-step kernels only, 152 + 186.7*k B/cell, k=2", "peak_source": peak_src},
        "cpu_baseline": cpu,
        "e2e": {"value": N / t_e2e / 1e6, "unit": "MLUP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": t_e2e * 1e3, "calls": KE, "ms_min": min(t_calls) * 1e3, "ms_max": max(t_calls) * 1e3,
                "api": "draw_circles + set_sinks (host lists in), ubgl_sim_step_host (fields out), tracers/items "
                       "advect, ubgl_items_download (item records out); no accumulator upload: items scatter on the device"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "kernels_ms_per_step": [{"kernel": k, "level": l, "launches": n, "ms": round(ms, 4)} for ms, n, k, l in kern[:14]],
    }
    print(json.dumps(line), flush=True)


def explosion_cpu_baseline(name, n_items=50_000, n_tracers=50_000):
    """The reference's CPU code for the same frame on a bounded sample: the fluid
    step (unmodified reference, all cores), its own advectFloatingItemsSimple on
    n_items, and the C restatement of the tracer shader on n_tracers."""
    from oracle import bind as ob
    from tests import next_cases
    if not ob.have_port():
        ob.build(ref=False)
    port = ob.Port()
    chk, kind = (ob.Ref(), "reference") if ob.have_ref() else (port, "port")
    nproc = chk.num_procs()
    chk.set_threads(nproc)
    W, H, flag, vx, vy, dt, items, st = explosion_inputs(name, max(n_items, n_tracers))
    sim = chk.Sim(flag, PWIDTH, MU)
    sim.set(ob.VX, vx)
    sim.set(ob.VY, vy)
    sim.step(dt)
    g = cases.LCG(99)
    full, simres = flag.copy(), flag.copy()
    t0 = time.perf_counter()
    for cx, cy, d in crater_list(g, W, H):
        port.draw_circle(full, simres, cx, cy, int(d), 1.0)
    sim.update_flag(simres)
    t_terrain = time.perf_counter() - t0
    t0 = time.perf_counter()
    sim.step(dt)
    t_step = time.perf_counter() - t0
    it = np.ascontiguousarray(items[:n_items])
    t0 = time.perf_counter()
    if kind == "reference":
        chk.items_advect_simple(sim, it, GAME_DT)
    else:
        ax, ay = np.zeros((H, W - 1), np.float32), np.zeros((H - 1, W), np.float32)
        port.items_advect_simple(it, GAME_DT, simres, sim.get(ob.VX), sim.get(ob.VY), sim.get(ob.P), ax, ay)
    t_items = time.perf_counter() - t0
    # tracers: crop of the co-located texture is not possible (positions span the domain): build it once
    S = 1024  # the tracer shader is timed against a 1024^2 member of the same generator (texture build is O(N))
    fl2, _ = cases.channel_flag(S, S, seed=1234)
    vx2, vy2 = cases.uniform_stream(fl2)
    vxy, _ = port.colocate(vx2, vy2)
    st2 = {k: np.ascontiguousarray(v[:n_tracers]) for k, v in st.items()}
    t0 = time.perf_counter()
    port.tracers_advect(st2, GAME_DT, (np.float32(PWIDTH), np.float32(PWIDTH)), 1, vxy, fl2)
    t_tr = time.perf_counter() - t0
    return {"value": W * H / t_step / 1e6, "unit": "MLUP/s", "cores": nproc, "kind": kind,
            "sample": f"{name}: 1 warm-up + 1 timed Simulation::step ({t_step * 1e3:.0f} ms, {nproc} OMP threads); "
                      f"16 craters + updateFields {t_terrain * 1e3:.1f} ms; advectFloatingItemsSimple on {n_items} items "
                      f"{t_items * 1e3:.0f} ms serial ({n_items / t_items:.3g} items/s; cost grows ~N^2/100); tracer shader "
                      f"restatement on {n_tracers} tracers {t_tr * 1e3:.1f} ms serial ({n_tracers / t_tr:.3g} tracers/s, port)",
            "items_per_s": n_items / t_items, "tracers_per_s": n_tracers / t_tr}


# algorithmic bytes per cell OF ITS LEVEL for each kernel kind (SURVEY.md 8d rule:
# every distinct input array once, every output once, fp32)
KERNEL_BYTES = {
    "rbgs_half": 8.0,           # half of a 16 B/cell red+black sweep
    "residual": 16.0, "restrict": 5.0, "prolong_correct": 22.0,
    "mg_pre_fused": 3 * 16.0 + 16.0 + 5.0,          # 3 sweeps + residual + restrict
    "mg_post_fused": 1.0 + 22.0 + 3 * 16.0,         # zero ec + prolong+correct + 3 sweeps
    "diffuse": 12.0, "accum": 16.0, "advect": 10.0, "divergence": 12.0, "gradient": 24.0,
    "mg_coarse_fused": 140.0 * 4.0 / 3.0,           # k_mg_tail: levels t..L, per cell of level t (geometric sum)
    "prestep_fused": 32.0 + 48.0, "advect_div_fused": 20.0 + 12.0, "finish_fused": 24.0 + 16.0,
}


# Kinds whose launch count per step depends on the schedule (advect: one merged launch for both
# components or one per component; prestep: per component a register-run launch for the interior plus
# a frame launch, which together cover the grid once): the figure is PER STEP, split over the launches.
KERNEL_BYTES_PER_STEP = {"advect": 2 * 10.0, "prestep_fused": 32.0 + 48.0}


def kernel_bytes_per_cell(kind, launches_per_step):
    """Algorithmic B per cell of its level for ONE launch of this kind."""
    if kind in KERNEL_BYTES_PER_STEP:
        return KERNEL_BYTES_PER_STEP[kind] / max(launches_per_step, 1)
    return KERNEL_BYTES.get(kind)


def ncu_traffic(workload, kind, level):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel from the
    committed `ncu --set full` capture of the same workload (profiles/), else None."""
    p = os.path.join(ROOT, "profiles", f"traffic_{workload}.json")
    if not workload or not os.path.exists(p):
        return None
    return json.load(open(p)).get("traffic_bytes_per_launch", {}).get(f"{kind}:{level}")


def traffic_source(workload):
    p = os.path.join(ROOT, "profiles", f"traffic_{workload}.json")
    if not workload or not os.path.exists(p):
        return None
    return f"profiles/traffic_{workload}.json (ncu --set full, DRAM read+write bytes per launch)"


def dominant_roofline(kern, W, H, peak, peak_src, prof_ms_step, cells_scale=1.0, workload=None):
    """`roofline` of the bench line: the (kernel, MG level) with the largest share of the step.
    achieved = algorithmic bytes per launch (KERNEL_BYTES x cells of its level) / mean launch ms."""
    if not kern:
        return None
    ms, n, k, l = kern[0]
    cells = (W >> l) * (H >> l) * cells_scale  # slab runs: one rank's rows
    bpc = kernel_bytes_per_cell(k, n)
    if bpc is None:
        return {"bound": "hbm", "kernel": k, "level": l, "achieved": None, "peak": peak, "unit": "GB/s",
                "frac": None, "traffic": None}
    per_launch_ms = ms / max(n, 1)
    bytes_launch = cells * bpc
    ach = bytes_launch / (per_launch_ms * 1e-3) / 1e9
    traffic = ncu_traffic(workload, k, l) if cells_scale == 1.0 else None
    out = {"bound": "hbm", "kernel": k, "level": l, "launches_per_step": n,
           "avg_launch_ms": per_launch_ms, "share_of_step": ms / prof_ms_step if prof_ms_step else None,
           "algorithmic_bytes_per_launch": bytes_launch, "achieved": ach, "peak": peak, "unit": "GB/s",
           "frac": ach / peak, "traffic": traffic, "peak_source": peak_src}
    if traffic is not None:
        out["traffic_source"] = traffic_source(workload)
        out["dram_frac"] = traffic / (per_launch_ms * 1e-3) / 1e9 / peak
    return out


def roofline_by_kernel(kern, W, H, peak, prof_ms_step, workload=None):
    """Every kernel kind summed over the MG levels it runs on (k_mg_run PRE + POST over the whole
    pyramid is the largest item of the step, no single (kernel, level) shows that).  `frac` is
    the algorithmic-byte roofline fraction; `dram_frac` = bytes ncu measured / time / peak, from
    the committed capture of the same workload (None when there is none)."""
    agg = {}
    for ms, n, k, l in kern:
        a = agg.setdefault(k, {"ms": 0.0, "launches": 0, "alg": 0.0, "traffic": 0.0, "have_traffic": True})
        a["ms"] += ms
        a["launches"] += n
        bpc = kernel_bytes_per_cell(k, n)
        if bpc is not None:
            a["alg"] += (W >> l) * (H >> l) * bpc * n
        t = ncu_traffic(workload, k, l)
        if t is None:
            a["have_traffic"] = False
        else:
            a["traffic"] += t * n
    rows = []
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        if a["ms"] <= 0:
            continue
        row = {"kernel": k, "launches_per_step": a["launches"], "ms_per_step": round(a["ms"], 4),
               "share_of_step": round(a["ms"] / prof_ms_step, 4) if prof_ms_step else None,
               "algorithmic_gbs": round(a["alg"] / (a["ms"] * 1e-3) / 1e9, 1) if a["alg"] else None,
               "frac": round(a["alg"] / (a["ms"] * 1e-3) / 1e9 / peak, 4) if a["alg"] else None,
               "dram_gbs": None, "dram_frac": None}
        if a["have_traffic"] and a["traffic"] > 0:
            row["dram_gbs"] = round(a["traffic"] / (a["ms"] * 1e-3) / 1e9, 1)
            row["dram_frac"] = round(row["dram_gbs"] / peak, 4)
        rows.append(row)
    return rows


def strong_scaling_base(name, dev, steps, warmup):
    """One GPU, the whole `name` grid resident (32768^2: 52 GB of state), single-GPU kernels."""
    import torch
    import ubootgl_b200 as u
    from ubootgl_b200 import capi
    try:
        W, H, flag, vx, vy, dt = make_inputs(name)
        sim = u.Simulation(flag, PWIDTH, MU, device=dev)
        sim.set(capi.VX, vx)
        del vx, vy, flag  # vy = 0 is the constructor state
        stream = torch.cuda.ExternalStream(sim.stream(), device=dev)
        for _ in range(warmup):
            sim.step(dt)
        sim.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            sim.step(dt)
        e1.record(stream)
        sim.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res = sim.residual()
        sim.close()
        return {"workload": name, "grid": [W, H], "n_gpus": 1, "ms_per_step": ms, "steps": steps, "warmup": warmup,
                "value": W * H / (ms * 1e-3) / 1e6, "unit": "MLUP/s", "residual_after": res,
                "note": "the N > 1 lines of bench.py --gpus N time this same workload; speed-up at N GPUs = "
                        "this ms_per_step / theirs"}
    except Exception as e:  # the headline number must survive a failure of the extra
        return {"workload": name, "error": f"{type(e).__name__}: {e}"[:300]}


def main():
    if len(sys.argv) >= 6 and sys.argv[1] == "--cpu-worker":  # child of cpu_reference_isolated
        t = int(sys.argv[5])
        print(json.dumps(cpu_reference_run(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), threads=t or None)),
              flush=True)
        return
    if len(sys.argv) >= 3 and sys.argv[1] == "--cpu-worker-explosion":
        print(json.dumps(explosion_cpu_baseline(sys.argv[2])), flush=True)
        return
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong-base", action="store_true",
                    help="skip the 1-GPU channel32768 run that anchors the strong-scaling study")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus == 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        if (args.workload or "").startswith("explosion"):
            return run_explosion(args, args.workload)
        return run_single_gpu(args, args.workload or "channel8192")
    from ubootgl_b200 import slab_bench, slab_boot
    try:
        return slab_bench.run(args, args.workload or "channel32768")
    finally:
        slab_boot.shutdown(barrier=False)  # no collective here: rank 0 is still busy with the CPU baseline


if __name__ == "__main__":
    main()
