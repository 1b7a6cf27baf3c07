"""Row-slab decomposition.  CPU: the plan arithmetic and the
two-process bootstrap over gloo.  GPU: the slab driver on one rank equals the
single-GPU Simulation bit for bit; with >= 2 GPUs the torchrun script
tests/mgpu_equiv.py checks 2 (and 4) ranks against one."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def plans(W, H, n):
    import ubootgl_b200 as u
    return [u.slab_plan(W, H, n, r) for r in range(n)]


@pytest.mark.parametrize("W,H,n", [(32768, 32768, 8), (32768, 32768, 4), (32768, 32768, 2),
                                   (8192, 8192, 8), (1090, 436, 2), (1000, 1536, 2), (1000, 1537, 4),
                                   (4096, 4096, 3), (512, 384, 2)])
def test_plan_covers_grid_and_aligns_cuts(W, H, n):
    ps = plans(W, H, n)
    nd = ps[0]["dist_levels"]
    assert nd >= 1 and all(p["dist_levels"] == nd for p in ps)
    assert ps[0]["own_lo"] == 0 and ps[-1]["own_hi"] == H
    for a, b in zip(ps, ps[1:]):
        assert a["own_hi"] == b["own_lo"]
    for p in ps:
        # cuts are multiples of 2^dist_levels: coarse row yc <-> fine row 2yc never straddles a cut
        assert p["own_lo"] % (1 << nd) == 0
        assert p["st_lo"] == max(0, p["own_lo"] - p["ghost"])
        assert p["st_hi"] == min(H, p["own_hi"] + p["ghost"])
        # every rank keeps at least two halos of rows on the coarsest distributed level
        assert ((p["own_hi"] - p["own_lo"]) >> (nd - 1)) >= 2 * p["ghost"]


def test_plan_rejects_too_many_ranks():
    import ubootgl_b200 as u
    with pytest.raises(u.UbglError):
        u.slab_plan(256, 256, 8, 0)
    with pytest.raises(u.UbglError):
        u.slab_plan(4096, 4096, 9, 0)


def test_flag_rows_generator_matches_global():
    f, _ = cases.channel_flag(640, 480, seed=1234)
    for p in plans(640, 480, 2):
        rows = cases.channel_flag_rows(640, 480, p["st_lo"], p["st_hi"], seed=1234)
        assert (rows == f[p["st_lo"]:p["st_hi"]]).all()


def test_weighted_plan_balances_fluid_rows():
    """ubgl_slab_set_row_weights: cuts at equal weight, still aligned, ordered, covering; reset restores."""
    import ubootgl_b200 as u
    W = H = 32768
    base = [u.slab_plan(W, H, 8, r) for r in range(8)]
    frac = cases.channel_row_fluid_fraction(W, H, seed=1234)
    wgt = (0.73 + 0.27 * frac).astype(np.float32)
    try:
        u.slab_set_row_weights(wgt)
        plans = [u.slab_plan(W, H, 8, r) for r in range(8)]
    finally:
        u.slab_set_row_weights(None)
    assert [u.slab_plan(W, H, 8, r) for r in range(8)] == base
    align = 1 << plans[0]["dist_levels"]
    assert plans[0]["own_lo"] == 0 and plans[-1]["own_hi"] == H
    for a, b in zip(plans, plans[1:]):
        assert a["own_hi"] == b["own_lo"] and a["own_hi"] % align == 0
    cost = [float(wgt[p["own_lo"]:p["own_hi"]].sum()) for p in plans]
    cost0 = [float(wgt[p["own_lo"]:p["own_hi"]].sum()) for p in base]
    assert max(cost) / min(cost) < max(cost0) / min(cost0)       # better balanced than equal heights
    assert max(cost) / (sum(cost) / 8) < 1.0 + 1.5 * align * float(wgt.max()) / (sum(cost) / 8)
    # the obstacle-free edge ranks got fewer rows than the middle ones
    rows = [p["own_hi"] - p["own_lo"] for p in plans]
    assert rows[0] < rows[3] and rows[7] < rows[4]


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
from ubootgl_b200 import slab_boot
import ubootgl_b200 as u
rank, world = slab_boot.init_distributed("gloo")
plan = u.slab_plan(4096, 4096, world, rank)
blob = bytes([rank]) * 64
blobs = slab_boot.blob_exchange()(blob)
assert [b[0] for b in blobs] == list(range(world)) and all(len(b) == 64 for b in blobs)
rows = np.full((plan["own_hi"] - plan["own_lo"], 5), float(rank + 1), np.float32)
full = slab_boot.gather_rows(plan["own_lo"], rows, 4096)
assert full.shape == (4096, 5) and (full[:2048] == 1).all() and (full[2048:] == 2).all()
assert slab_boot.allreduce_max(rank) == world - 1
assert slab_boot.allreduce_sum(rank + 1) == world * (world + 1) / 2
slab_boot.barrier()
print("GLOO_OK", rank, flush=True)
slab_boot.shutdown()
"""


def test_bootstrap_two_ranks_gloo(tmp_path):
    """world_size 2 over gloo on CPU: IPC-blob all-gather, row gather, reductions."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    import socket
    with socket.socket() as so:  # a free port: a fixed one collides with a lingering earlier run
        so.bind(("127.0.0.1", 0))
        port = so.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"GLOO_OK {r}" in o, o


def bits_same(a, b):
    return bool(((a.view(np.uint32) == b.view(np.uint32)) | ((a == 0) & (b == 0))).all())


@pytest.mark.gpu
@pytest.mark.parametrize("W,H", [(258, 200), (600, 333)])
def test_one_rank_slab_equals_simulation(ubgl, W, H):
    from ubootgl_b200 import capi
    c = cases.sim_case(W, H, seed=W + H)
    S = ubgl.SlabSimulation(c["flag"], W, H, 0, 1, lambda b: [b], device=0)
    G = ubgl.Simulation(c["flag"])
    for s in (S, G):
        put = s.set_from_global if s is S else s.set
        put(capi.VX, c["vx"]); put(capi.VY, c["vy"])
        put(capi.VX_ACCUM, c["vx_accum"]); put(capi.VY_ACCUM, c["vy_accum"]); put(capi.P, c["p"])
        s.set_sinks([[0.4, 0.4 * H / W, 120.0]])
    for _ in range(3):
        S.step(0.001)
        G.step(0.001)
    S.sync()
    for fld in (capi.VX, capi.VY, capi.P, capi.F, capi.VXB, capi.VYB, capi.VX_CURRENT, capi.VY_CURRENT,
                capi.VX_ACCUM):
        assert bits_same(S.get(fld)[1], G.get(fld)), fld
    assert abs(np.sqrt(S.residual_sumsq()) - G.residual()) <= 1e-5 * G.residual()


@pytest.mark.gpu
def test_one_rank_slab_step_host_equals_simulation_step_host(ubgl):
    """ubgl_slab_step_host (accumulator mirrors in, vx / vy / p / *_current mirrors out, packed DMAs)
    against ubgl_sim_step_host, bit for bit, incl. the clearing of the accumulator mirrors."""
    from ubootgl_b200 import capi
    W, H = 600, 333
    c = cases.sim_case(W, H, seed=77)
    S = ubgl.SlabSimulation(c["flag"], W, H, 0, 1, lambda b: [b], device=0)
    G = ubgl.Simulation(c["flag"])
    for s in (S, G):
        put = s.set_from_global if s is S else s.set
        put(capi.VX, c["vx"]); put(capi.VY, c["vy"]); put(capi.P, c["p"])
    shapes = dict(vx_accum=(H, W - 1), vy_accum=(H - 1, W), vx=(H, W - 1), vy=(H - 1, W), p=(H, W),
                  vx_current=(H, W - 1), vy_current=(H - 1, W))
    for k in range(2):
        bufs = []
        for s in (S, G):
            b = {n: np.full(sh, np.nan, np.float32) for n, sh in shapes.items()}
            b["vx_accum"][:] = c["vx_accum"] * (k + 1)
            b["vy_accum"][:] = c["vy_accum"] * (k + 1)
            s.step_host(0.001, **b)
            bufs.append(b)
        for n in shapes:
            assert bits_same(bufs[0][n], bufs[1][n]), (k, n)
        assert not bufs[0]["vx_accum"][1:-1, 1:-1].any() and bufs[0]["vx_accum"][0].any()  # interior cleared, border kept
        assert (bufs[0]["vx"] == bufs[0]["vx_current"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("nranks,dt", [(2, "0.002"), (2, "0.02"), (4, "0.002"), (4, "0.02"), (2, "0.002 skew"), (4, "0.02 skew")])
def test_slabs_equal_single_gpu(ubgl, nranks, dt):
    if ubgl.lib.ubgl_device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs (run under gpurun --gpus {nranks})")
    cmd = ["timeout", "600", sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1", "--master-port", str(29540 + nranks),
           os.path.join(ROOT, "tests", "mgpu_equiv.py"), "1000", "1536", "3"] + dt.split()
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and "MGPU_EQUIV OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
