"""CPU: bench.py's reference arm (`--impl reference`) and the isolated CPU-baseline runner print the
contract's JSON line without a GPU (the driver runs this arm on the GPU box's host cores)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "channel256", "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "fluid_step_throughput" and line["unit"] == "MLUP/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["config"]["workload"] == "channel256" and line["config"]["grid"] == [256, 256]
    # what is printed is what ran (round-1 printed the requested 20 steps for 3 timed ones)
    assert line["steps"] == 2 and line["warmup"] == 1 and line["config"]["sampled_for"] is None
    # both arms describe the workload with the same `config` object
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.workload_config("channel256")
    cb, e2e = line["cpu_baseline"], line["e2e"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_isolated_cpu_runner_survives_and_reports():
    sys.path.insert(0, ROOT)
    import bench
    r = bench.cpu_reference_isolated("channel256", 1, 1, threads=1)
    assert r is not None and r["cores"] == 1 and r["W"] == 256 and r["ms_per_step"] > 0
    # a child that dies (unknown workload -> SystemExit) is reported as None, not as an exception
    assert bench.cpu_reference_isolated("no-such-workload", 1, 1) is None
    assert bench.pcie_d2h_bytes({"vx": __import__("numpy").zeros(4, "f4"), "vx_current": __import__("numpy").zeros(4, "f4")}) == 32


def test_reference_arm_non_zero_ranks_exit_quietly():
    """Under torchrun (N > 1) only rank 0 runs the reference; the others exit 0 without output."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT,
                         env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_labels_a_sampled_workload():
    """channel32768 cannot be stepped by the CPU reference inside a bench run: config.workload says what ran
    (channel8192) and config.sampled_for what it stands in for; keys equal our arm's."""
    sys.path.insert(0, ROOT)
    import bench
    c = bench.workload_config("channel8192", "channel32768", parallelism="row slabs x8")
    assert c["workload"] == "channel8192" and c["grid"] == [8192, 8192] and c["sampled_for"] == "channel32768"
    assert set(c) == set(bench.workload_config("channel32768", parallelism="row slabs x8"))


def test_roofline_by_kernel_aggregates_levels():
    sys.path.insert(0, ROOT)
    import bench
    kern = [(0.9, 2, "mg_pre_fused", 0), (0.8, 2, "mg_post_fused", 0), (0.23, 2, "mg_pre_fused", 1), (1.1, 1, "advect", 0)]
    rows = bench.roofline_by_kernel(kern, 8192, 8192, 6548.8, 3.03, workload="no-such-capture")
    by = {r["kernel"]: r for r in rows}
    assert abs(by["mg_pre_fused"]["ms_per_step"] - 1.13) < 1e-9 and by["mg_pre_fused"]["launches_per_step"] == 4
    alg = (8192 * 8192 * 69.0 * 2 + 4096 * 4096 * 69.0 * 2) / 1.13e-3 / 1e9
    assert abs(by["mg_pre_fused"]["algorithmic_gbs"] - alg) < 1.0
    assert by["advect"]["dram_frac"] is None  # no committed ncu capture for this workload: no phantom source
    r = bench.dominant_roofline(sorted(kern, reverse=True), 8192, 8192, 6548.8, "measured", 3.03, workload="no-such-capture")
    assert r["traffic"] is None and "traffic_source" not in r
