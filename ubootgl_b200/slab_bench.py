"""bench.py --gpus N (N > 1): the fluid step row-slab decomposed over N B200s of
one box, one process per GPU (torchrun).  Same JSON contract as the 1-GPU arm;
timing = CUDA events on each rank's library stream between two barriers, MAX
over ranks."""
import json
import os
import time

import numpy as np


def run(args, name):
    import torch
    import ubootgl_b200 as u
    from ubootgl_b200 import capi, slab_boot
    import bench
    from tests import cases

    rank, world = slab_boot.init_distributed("nccl")
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    if world != args.gpus:
        raise SystemExit(f"bench.py --gpus {args.gpus} launched with WORLD_SIZE={world}")
    W, H = bench.workload_dims(name)
    N = W * H

    # ---- N-GPU == 1-GPU, bit for bit, BEFORE anything is timed: the slab decomposition of a
    # seeded 1000 x 1536 (8 ranks: x 3072) case on all ranks against the single-GPU Simulation on rank 0, at
    # CFL ~ 1-3 (back-traces stay in the ghost rows) and CFL ~ 25-75 (taps served by NVLink
    # peer loads).  Every scaling number below rests on this equivalence; a mismatch aborts. ----
    from tests import mgpu_equiv
    eh = max(1536, 384 * world)  # >= 384 rows per rank: the CFL ~ 75 back-traces stay inside the neighbouring slab
    equiv_runs = [mgpu_equiv.check(1000, eh, 2, dt_, rank, world, dev, verbose=(rank == 0), skew=sk)
                  for dt_, sk in ((0.002, False), (0.02, True))]  # the second with unequal (weighted) slab heights
    equiv = {"bitwise_ok": all(r["bitwise_ok"] for r in equiv_runs), "fields": equiv_runs[0]["fields"],
             "ranks": world, "grid": equiv_runs[0]["grid"], "steps": 2, "dt": [r["dt"] for r in equiv_runs],
             "dist_levels": equiv_runs[0]["dist_levels"], "exchanges": [r["exchanges"] for r in equiv_runs],
             "residual": [r["residual"] for r in equiv_runs],
             "mismatched_fields": [r["mismatched_fields"] for r in equiv_runs],
             "what": "slab-decomposed step on all ranks vs one-GPU Simulation on rank 0: vx, vy, p, vx_current, "
                     "vy_current, vx back buffer, f, vx_accum compared as uint32 (sign of zero ignored) + residual norm"}
    if not equiv["bitwise_ok"]:
        if rank == 0:
            print(json.dumps({"metric": "fluid_step_throughput", "n_gpus": world, "error": "slab != single GPU",
                              "equiv": equiv}), flush=True)
        raise SystemExit(3)

    # load balance: the advect pass (27 % of a step) skips octets without fluid, every other pass
    # costs the same per row; cuts at equal weight instead of equal height (identical on every
    # rank: computed from the generator's disc list)
    # (measured at 8 GPUs, m2: 11.46 ms weighted vs 11.29 ms equal heights -- no gain, the exchanges wait on
    # jitter, not on a systematic imbalance; the API stays, the bench uses equal heights unless asked)
    if os.environ.get("UBGL_SLAB_BALANCE", "0") == "1":
        frac = cases.channel_row_fluid_fraction(W, H, seed=1234)
        # fit of the per-rank advect times of an 8-GPU profile (gpurun_out/s2_tl.*): t = 0.739 + 2.063 frac ms per
        # 4096 rows, beside 6.4 ms per 4096 rows for everything else
        u.slab_set_row_weights((0.776 + 0.224 * frac).astype(np.float32))
    dt = float(np.float32(bench.PWIDTH) / np.float32(W - 1))  # dt = h, CFL ~ 1
    if os.environ.get("UBGL_SLAB_SWEEP_MIN_ROWS"):  # tuning run: distributed-level depth (slab.cu: slab_min_rows)
        import sys
        for mr in os.environ["UBGL_SLAB_SWEEP_MIN_ROWS"].split(","):
            os.environ["UBGL_SLAB_MIN_ROWS"] = mr
            pl = u.slab_plan(W, H, world, rank)
            fl = cases.channel_flag_rows(W, H, pl["st_lo"], pl["st_hi"], seed=1234)
            sm = u.SlabSimulation(fl, W, H, rank, world, slab_boot.blob_exchange(), bench.PWIDTH, bench.MU, device=dev)
            v0 = (fl[:, :-1] * fl[:, 1:]).astype(np.float32)
            v0[:, 0] = 1.0
            sm.set(capi.VX, v0)
            st = torch.cuda.ExternalStream(sm.stream(), device=dev)
            for _ in range(3):
                sm.step(dt)
            sm.sync()
            blocks = []
            for _ in range(3):  # three timed blocks of 10 steps
                slab_boot.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st)
                for _ in range(10):
                    sm.step(dt)
                b.record(st)
                sm.sync()
                blocks.append(round(slab_boot.allreduce_max(a.elapsed_time(b)) / 10, 3))
            x_, _b = sm.stats()
            res = float(np.sqrt(slab_boot.allreduce_sum(sm.residual_sumsq())))
            if rank == 0:
                print(f"[sweep] min_rows {mr}: dist levels {pl['dist_levels']}, ms/step per block {blocks}, "
                      f"exchanges/step {x_ / 33:.1f}, residual {res:.6f}", file=sys.stderr, flush=True)
            slab_boot.barrier()
            sm.close()
            del sm, st
            slab_boot.barrier()
        os.environ.pop("UBGL_SLAB_MIN_ROWS", None)
    plan = u.slab_plan(W, H, world, rank)
    all_rows = [u.slab_plan(W, H, world, r) for r in range(world)]
    # synthetic input, generated slab-wise: only the rows this rank stores
    flag = cases.channel_flag_rows(W, H, plan["st_lo"], plan["st_hi"], seed=1234)
    sim = u.SlabSimulation(flag, W, H, rank, world, slab_boot.blob_exchange(), bench.PWIDTH, bench.MU,
                           device=dev)
    vx = (flag[:, :-1] * flag[:, 1:]).astype(np.float32)
    vx[:, 0] = 1.0
    sim.set(capi.VX, vx)
    del vx
    stream = torch.cuda.ExternalStream(sim.stream(), device=dev)
    K, Wm = args.steps, args.warmup

    for _ in range(Wm):
        sim.step(dt)
    sim.sync()
    torch.cuda.synchronize()
    slab_boot.barrier()
    l0 = sim.launch_count()
    x0, b0 = sim.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # Clocks and throttle reasons DURING the timed region, without a sampling thread: the ranks run in
    # lockstep (34 neighbour exchanges per step), so an NVML query that holds up ONE rank's launches for a
    # millisecond holds up all of them -- a 50 ms sampling thread per rank cost 0.65 ms per step at 8 GPUs
    # (10.0 vs 9.35 ms, tools/gpu_t.sh).  The launches of the K steps are queued well ahead of the GPU; every
    # rank then samples from this thread while its GPU is still working through them.
    clk = bench.ClockSampler(dev)
    e0.record(stream)
    for _ in range(K):
        sim.step(dt)
    e1.record(stream)
    for _ in range(3):
        if e1.query():
            break
        clk.sample_once()
        time.sleep(0.01)
    sim.sync()
    torch.cuda.synchronize()
    if not clk.rows:  # a run too short to catch in flight: one sample right behind it
        clk.sample_once()
    slab_boot.barrier()
    ms_total = slab_boot.allreduce_max(e0.elapsed_time(e1))
    launches = slab_boot.allreduce_sum(sim.launch_count() - l0)
    x1, b1 = sim.stats()
    halo_mb = slab_boot.allreduce_sum((b1 - b0) / K / 1e6)
    ms_step = ms_total / K
    value = N * K / (ms_total * 1e-3) / 1e6
    res_after = float(np.sqrt(slab_boot.allreduce_sum(sim.residual_sumsq())))

    # ---- per-kernel profile of every rank (CUDA events around each launch) ----
    sim.profile(True)
    PK = min(K, 3)
    for _ in range(PK):
        sim.step(dt)
    sim.sync()
    stats = sim.kernel_stats()
    sim.profile(False)
    prof_total = sum(ms for _, ms in stats.values()) / PK
    kern = sorted(((ms / PK, n // PK, k, l) for (k, l), (n, ms) in stats.items()), reverse=True)
    wait_ms = sum(ms for ms, n, k, l in kern if k == "halo_wait")
    push_ms = sum(ms for ms, n, k, l in kern if k == "halo_push")
    wait_max, wait_min = slab_boot.allreduce_max(wait_ms), -slab_boot.allreduce_max(-wait_ms)
    push_max = slab_boot.allreduce_max(push_ms)
    compute_max = slab_boot.allreduce_max(prof_total - wait_ms - push_ms)
    compute_min = -slab_boot.allreduce_max(-(prof_total - wait_ms - push_ms))

    # ---- end to end: each rank feeds its slab's accumulators from pinned host memory and reads
    # its rows of vx, vy, p, vx_current, vy_current back (ubgl_slab_step_host: contiguous DMAs of
    # device-packed rows; the *_current mirrors are host copies of the landed vx / vy bands) ----
    pin = lambda shape: torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
    names = dict(vx_accum=capi.VX_ACCUM, vy_accum=capi.VY_ACCUM, vx=capi.VX, vy=capi.VY, p=capi.P,
                 vx_current=capi.VX_CURRENT, vy_current=capi.VY_CURRENT)
    bufs = {k: pin(sim.field_rows(f)[1:]) for k, f in names.items()}
    bufs["vx_accum"][:] = 0
    bufs["vy_accum"][:] = 0
    own = plan["own_hi"] - plan["own_lo"]
    h2d = bufs["vx_accum"].nbytes + bufs["vy_accum"].nbytes
    d2h = sum(min(own, bufs[k].shape[0]) * bufs[k].shape[1] * 4 for k in ("vx", "vy", "p"))  # one crossing each

    def host_step():
        sim.step_host(dt, **bufs)

    KE = max(1, min(K, 2))
    host_step()
    slab_boot.barrier()
    t0 = time.perf_counter()
    for _ in range(KE):
        host_step()
    slab_boot.barrier()
    t_e2e = slab_boot.allreduce_max((time.perf_counter() - t0) / KE)
    h2d_all, d2h_all = slab_boot.allreduce_sum(h2d), slab_boot.allreduce_sum(d2h)
    if os.environ.get("UBGL_SLAB_E2E_SWEEP"):  # diagnosis: host copy threads per rank, and without the *_current mirrors
        sweep = {}
        for nt in (16, 4, 2, 1):
            os.environ["UBGL_HOST_THREADS"] = str(nt)
            for tag, bb in (("all", bufs), ("no_current", {k: v for k, v in bufs.items() if "current" not in k}),
                            ("no_accum", {k: v for k, v in bufs.items() if "accum" not in k})):
                if nt != 16 and tag != "all":
                    continue
                sim.step_host(dt, **bb)
                slab_boot.barrier()
                t0 = time.perf_counter()
                sim.step_host(dt, **bb)
                mine = time.perf_counter() - t0
                slab_boot.barrier()
                sweep[f"{tag}_threads{nt}"] = (round(slab_boot.allreduce_max(mine) * 1e3, 1),
                                               round(-slab_boot.allreduce_max(-mine) * 1e3, 1))
        os.environ.pop("UBGL_HOST_THREADS", None)
        if rank == 0:
            print("e2e sweep, ms (max rank, min rank):", json.dumps(sweep), flush=True)

    clocks = clk.summary()
    # the slowest clock and the union of the throttle reasons over all ranks
    bits = sum(1 << i for i, nm in enumerate(bench.ClockSampler.NAMES) if nm in clocks["reasons"])
    allbits = 0
    for i in range(len(bench.ClockSampler.NAMES)):
        if slab_boot.allreduce_max(float((bits >> i) & 1)) > 0:
            allbits |= 1 << i
    clocks["reasons"] = [nm for i, nm in enumerate(bench.ClockSampler.NAMES) if allbits & (1 << i)]
    lo = -slab_boot.allreduce_max(-(clocks["sm_mhz"] if clocks["sm_mhz"] is not None else 1e9))
    clocks["sm_mhz"] = lo if lo < 1e9 else None
    clocks["samples"] = int(slab_boot.allreduce_sum(float(clocks["samples"])))
    clocks["source"] += " (every rank, from the launching thread after the steps were queued; min clock / union of reasons over ranks)"
    if rank != 0:
        return
    cpu = None
    if not args.no_cpu_baseline:  # rank 0 only; the other ranks are done
        cpu = bench.cpu_baseline_entry("channel8192", steps=3, warmup=1, policy=False)
        if cpu.get("sample"):
            cpu["sample"] += f" -- sampled for {name} (a 32768^2 reference step needs ~80 GB of host memory)"
    peak, peak_src = bench.peaks()
    bpc = bench.bytes_per_cell()
    step_gbs = N * bpc / (ms_step * 1e-3) / 1e9
    line = {
        "metric": "fluid_step_throughput", "value": value, "unit": "MLUP/s", "n_gpus": world, "steps": K,
        "warmup": Wm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench.workload_config(name, parallelism=f"row slabs x{world}"),
        "run_info": {"residual_after": res_after, "decomposition": f"{plan['dist_levels']} distributed MG levels, "
                     f"coarser levels replicated, ghost {plan['ghost']} rows",
                     "rows_per_rank": [r["own_hi"] - r["own_lo"] for r in all_rows],
                     "halo_mb_per_step_all_ranks": halo_mb, "exchanges_per_step": (x1 - x0) / K,
                     "scaling_note": "strong scaling is defined on this workload; its 1-GPU base is the "
                                     "strong_scaling_base object of the N = 1 line (bench.py --gpus 1)"},
        "equiv": equiv,
        "roofline": bench.dominant_roofline([r for r in kern if not r[2].startswith("halo")], W, H, peak, peak_src,
                                            prof_total, cells_scale=1.0 / world),
        "roofline_step": {"bound": "hbm", "kernel": "whole step (all ranks)", "achieved": step_gbs,
                          "peak": peak * world, "unit": "GB/s", "frac": step_gbs / (peak * world), "traffic": None,
                          "bytes_per_cell": bpc, "peak_source": peak_src + f" x {world} GPUs",
                          "model": "stage-wise algorithmic bytes 152 + 186.7*k B/cell (SURVEY.md 8d), k=2"},
        "halo": {"wait_ms_per_step_max_rank": wait_max, "wait_ms_per_step_min_rank": wait_min,
                 "push_ms_per_step_max_rank": push_max, "compute_ms_per_step_max_rank": compute_max,
                 "compute_ms_per_step_min_rank": compute_min,
                 "note": "profiled run (events around every launch); one k_halo_push launch per exchange "
                         "stores the boundary rows into the neighbours, releases, then waits for their "
                         "releases: push time = NVLink stores + load imbalance + latency (wait is 0: "
                         "the separate k_halo_wait launch is gone)"},
        "kernels_ms_per_step_rank0": [{"kernel": k, "level": l, "launches": n, "ms": round(ms, 4)}
                                      for ms, n, k, l in kern[:14]],
        "cpu_baseline": cpu,
        "e2e": {"value": N / t_e2e / 1e6, "unit": "MLUP/s", "h2d_bytes_per_step": int(h2d_all),
                "d2h_bytes_per_step": int(d2h_all), "ms_per_step": t_e2e * 1e3,
                "api": "ubgl_slab_step_host on every rank (pinned host slabs): accumulators in; own rows of vx, vy, "
                       "p out over each rank's PCIe link, packed on the device; vx_current, vy_current filled "
                       "from the vx, vy mirrors by host threads like saveCurrentVelocityFields' memcpy"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
