"""GPU parity: every stage of Simulation::step and whole steps (libubgl.so
through the C ABI) against the CPU oracle on identical seeded inputs."""
import numpy as np
import pytest

from oracle import bind as ob
from tests import cases
from tests.cases import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-5
# widths chosen to hit every "last 0-7 columns untouched" residue (W mod 8) and
# the flat-index wrap of the vy loop (W = 2 mod 8, e.g. 1090)
SIZES = [(24, 16), (41, 33), (66, 50), (70, 40), (71, 40), (72, 41), (73, 40), (74, 44),
         (75, 40), (76, 40), (77, 40), (130, 97), (258, 131)]
FIELDS = [ob.VX, ob.VY, ob.VXB, ob.VYB, ob.P, ob.F, ob.VX_ACCUM, ob.VY_ACCUM]


def make_pair(ubgl, port, W, H, seed):
    c = cases.sim_case(W, H, seed)
    G = ubgl.Simulation(c["flag"])
    O = port.Sim(c["flag"])
    for s in (G, O):
        s.set(ob.VX, c["vx"])
        s.set(ob.VY, c["vy"])
        s.set(ob.VXB, c["vx"][::-1].copy())  # distinct junk in the back buffers
        s.set(ob.VYB, c["vy"][::-1].copy())
        s.set(ob.VX_ACCUM, c["vx_accum"])
        s.set(ob.VY_ACCUM, c["vy_accum"])
        s.set(ob.P, c["p"])
    return G, O, c


def sync_from_oracle(G, O):
    for f in FIELDS:
        G.set(f, O.get(f))


FIELD_NAMES = {ob.VX: "vx", ob.VY: "vy", ob.VXB: "vx_back", ob.VYB: "vy_back", ob.P: "p", ob.F: "f",
               ob.VX_ACCUM: "vx_accum", ob.VY_ACCUM: "vy_accum", ob.VX_CURRENT: "vx_current", ob.VY_CURRENT: "vy_current"}


def check(G, O, fields, tol=TOL, what=""):
    errs = {}
    for f in fields:
        a, b = G.get(f), O.get(f)
        assert a.shape == b.shape
        errs[FIELD_NAMES.get(f, str(f))] = rel_l2(a, b)
    # measured errors go to gpurun_out/parity_errors.jsonl (profiles/r02_parity_errors.md: worst per stage)
    cases.record_parity(f"test_gpu_sim: {what} (tol {tol:g})", O.get(ob.P).shape[::-1], "restatement", errs)
    for k, e in errs.items():
        assert e <= tol, (what, k, e)


@pytest.mark.parametrize("W,H", SIZES)
def test_stages(ubgl, port, W, H):
    G, O, c = make_pair(ubgl, port, W, H, seed=W * 100 + H)
    dt = float(O.dx)  # CFL ~ 1
    G.stage(ob.ST_SETVBCS, dt); O.stage(ob.ST_SETVBCS, dt)
    check(G, O, [ob.VX, ob.VY, ob.VXB, ob.VYB], 0.0, "setVBCs")
    G.stage(ob.ST_ACCUM, dt); O.stage(ob.ST_ACCUM, dt)
    check(G, O, [ob.VX, ob.VY, ob.VX_ACCUM, ob.VY_ACCUM], 1e-7, "accum")
    sync_from_oracle(G, O)
    G.stage(ob.ST_DIFFUSE, dt); O.stage(ob.ST_DIFFUSE, dt)
    check(G, O, [ob.VX, ob.VY, ob.VXB, ob.VYB], 2e-7, "diffuse")   # observed <= 7.1e-8 (profiles/r02_parity_errors.md)
    sync_from_oracle(G, O)
    G.stage(ob.ST_ADVECT, dt); O.stage(ob.ST_ADVECT, dt)
    check(G, O, [ob.VX, ob.VY, ob.VXB, ob.VYB], 3e-6, "advect")    # observed <= 1.5e-6
    # untouched entries (skipped octets, last columns) are bit-identical copies
    gx, ox = G.get(ob.VX), O.get(ob.VX)
    stale = ox == c["vx"][::-1] if False else None
    sync_from_oracle(G, O)
    G.stage(ob.ST_SETVBCS, dt); O.stage(ob.ST_SETVBCS, dt)
    check(G, O, [ob.VX, ob.VY, ob.VXB, ob.VYB], 0.0, "setVBCs2")
    G.stage(ob.ST_PROJECT, dt); O.stage(ob.ST_PROJECT, dt)
    check(G, O, [ob.F], 1e-7, "divergence")                         # observed 0
    check(G, O, [ob.P, ob.VX, ob.VY], 3e-6, "project")              # observed <= 1.5e-6
    sync_from_oracle(G, O)
    G.stage(ob.ST_SAVE, dt); O.stage(ob.ST_SAVE, dt)
    check(G, O, [ob.VX_CURRENT, ob.VY_CURRENT], 0.0, "save")


@pytest.mark.parametrize("W,H", [(70, 40), (74, 44), (130, 97)])
def test_advect_quirks_bit_exact(ubgl, port, W, H):
    """Faces the reference never advects (whole-octet skip, last 0-7 columns)
    must keep the back-buffer content bit for bit."""
    G, O, c = make_pair(ubgl, port, W, H, seed=11)
    dt = float(O.dx)
    backx, backy = O.get(ob.VXB), O.get(ob.VYB)
    G.stage(ob.ST_ADVECT, dt); O.stage(ob.ST_ADVECT, dt)
    gx, ox, gy, oy = G.get(ob.VX), O.get(ob.VX), G.get(ob.VY), O.get(ob.VY)
    sx = (ox.view(np.uint32) == backx.view(np.uint32))
    sy = (oy.view(np.uint32) == backy.view(np.uint32))
    assert sx.sum() > 0 and sy.sum() > 0
    assert (gx.view(np.uint32)[sx] == backx.view(np.uint32)[sx]).all()
    assert (gy.view(np.uint32)[sy] == backy.view(np.uint32)[sy]).all()
    # and the GPU leaves exactly the same set untouched (up to value coincidences)
    tx = (gx.view(np.uint32) == backx.view(np.uint32))
    assert abs(int(tx.sum()) - int(sx.sum())) <= 2


@pytest.mark.parametrize("bcs", [(0, 2, 3, 3), (3, 3, 3, 3), (0, 1, 3, 3), (1, 2, 0, 3), (2, 0, 1, 1)])
def test_boundary_condition_kinds(ubgl, port, bcs):
    G, O, c = make_pair(ubgl, port, 66, 50, seed=9)
    G.set_bc(*bcs); O.set_bc(*bcs)
    dt = float(O.dx)
    G.stage(ob.ST_SETVBCS, dt); O.stage(ob.ST_SETVBCS, dt)
    check(G, O, [ob.VX, ob.VY, ob.VXB, ob.VYB], 0.0, "setVBCs")
    G.step(dt); O.step(dt)
    check(G, O, [ob.VX, ob.VY, ob.P], 2e-6, "step")  # observed <= 8.7e-7


@pytest.mark.parametrize("W,H", [(70, 40), (130, 97), (258, 131), (1090, 436)])
def test_steps_with_sinks(ubgl, port, W, H):
    G, O, c = make_pair(ubgl, port, W, H, seed=W + H)
    for s in (G, O):
        s.add_sink(0.4, 0.4 * H / W, 120.0)
        s.add_sink(0.401, 0.4 * H / W, 60.0)   # overlapping stamp: list order wins
        s.add_sink(0.001, 0.001, 50.0)         # inside the 3-cell border: skipped, never decays
    dt = 0.001
    for k in range(3):
        G.step(dt); O.step(dt)
        tol = (7e-6, 1.3e-5, 3.2e-5)[k]  # observed <= 3.3e-6, 6.5e-6, 1.6e-5 (vy at 1090x436): twice that
        check(G, O, [ob.VX, ob.VY, ob.P, ob.VX_CURRENT, ob.VY_CURRENT], tol, f"step {k}")
        assert np.allclose(G.sinks(), O.sinks(), rtol=1e-6, atol=0)
        assert (G.get(ob.VX_ACCUM) == O.get(ob.VX_ACCUM)).all()


def test_step_host_mirrors(ubgl, port):
    """The e2e entry point: host accumulators in (and zeroed), fields out."""
    W, H = 130, 97
    G, O, c = make_pair(ubgl, port, W, H, seed=1)
    dt = 0.001
    ax, ay = c["vx_accum"].copy(), c["vy_accum"].copy()
    out = {k: np.empty(s, np.float32) for k, s in
           dict(vx=(H, W - 1), vy=(H - 1, W), p=(H, W), vx_current=(H, W - 1),
                vy_current=(H - 1, W)).items()}
    G.step_host(dt, flag=c["flag"], vx_accum=ax, vy_accum=ay, **out)
    O.step(dt)
    assert rel_l2(out["vx"], O.get(ob.VX)) <= 5e-6  # one step at 130x97: observed <= 1.6e-6
    assert rel_l2(out["vy"], O.get(ob.VY)) <= 5e-6
    assert rel_l2(out["p"], O.get(ob.P)) <= 5e-6
    assert (out["vx_current"] == out["vx"]).all() and (out["vy_current"] == out["vy"]).all()
    assert (ax == O.get(ob.VX_ACCUM)).all() and (ay == O.get(ob.VY_ACCUM)).all()


def test_step_host_banded_current_mirrors(ubgl):
    """Above 16 MB per field the velocity mirrors come down in row bands and the
    *_current mirrors are host copies of them (saveCurrentVelocityFields is a memcpy,
    simulation.cpp:16-19): every mirror must equal the device field bit for bit."""
    from ubootgl_b200 import capi
    W, H = 2307, 2050
    c = cases.sim_case(W, H, seed=11)
    G = ubgl.Simulation(c["flag"], device=0)
    G.set(capi.VX, c["vx"]); G.set(capi.VY, c["vy"])
    ax, ay = c["vx_accum"].copy(), c["vy_accum"].copy()
    out = {k: np.full(s, np.nan, np.float32) for k, s in
           dict(vx=(H, W - 1), vy=(H - 1, W), p=(H, W), vx_current=(H, W - 1),
                vy_current=(H - 1, W)).items()}
    for _ in range(2):
        G.step_host(0.0005, vx_accum=ax, vy_accum=ay, **out)
    for name, fid in (("vx", capi.VX), ("vy", capi.VY), ("p", capi.P),
                      ("vx_current", capi.VX_CURRENT), ("vy_current", capi.VY_CURRENT)):
        assert np.array_equal(out[name], G.get(fid)), name
    assert not ax[1:-1, 1:-2].any() and not ay[1:-2, 1:-1].any()
    # only one of the pair requested: plain downloads
    only = np.full((H, W - 1), np.nan, np.float32)
    G.step_host(0.0005, vx_current=only)
    assert np.array_equal(only, G.get(capi.VX_CURRENT))


def test_flag_update_rebuilds_pyramid(ubgl, port):
    W, H = 130, 97
    G, O, c = make_pair(ubgl, port, W, H, seed=2)
    flag2 = c["flag"].copy()
    flag2[40:60, 50:80] = 0
    G.update_flag(flag2); O.update_flag(flag2)
    assert G.mg_levels() == O.mg_levels()
    for l in range(G.mg_levels()):
        assert (G.mg_flagc(l) == O.mg_flagc(l)).all()
    G.step(0.001); O.step(0.001)
    check(G, O, [ob.VX, ob.VY, ob.P], 3.5e-6, "step after flag edit")  # observed <= 1.6e-6


def test_tolerance_mode_matches_fixed_count_and_oracle_history(ubgl, port):
    """Tolerance / stagnation stop rule of the pressure solves:
    the cycles it runs are the reference's V-cycles -- same fields as a fixed-count
    run of that many cycles, residual history within 1 % of the oracle's MG."""
    W, H = 258, 131
    G, O, c = make_pair(ubgl, port, W, H, seed=21)
    G2, _, _ = make_pair(ubgl, port, W, H, seed=21)
    dt = 0.001
    G.set_tolerance(1e-4, max_cycles=12, stagnation=0.9)
    G.step(dt)
    n, fnorm, hist = G.solve_info()
    assert 1 <= n <= 12 and len(hist) == n + 1 and fnorm > 0
    # stop rule: every cycle but the last made progress and did not reach the target
    for k in range(1, n):
        assert hist[k] > 1e-4 * fnorm and hist[k] < 0.9 * hist[k - 1]
    assert n == 12 or hist[n] <= 1e-4 * fnorm or hist[n] >= 0.9 * hist[n - 1]
    G2.set_option(ubgl.capi.OPT_VCYCLES, n)
    G2.step(dt)
    for f in (ob.VX, ob.VY, ob.P):
        assert (G.get(f) == G2.get(f)).all()
    # oracle: MG::solve on the same rhs, warm start = the p the step began with
    f = G.get(ob.F)
    fl = c["flag"]
    assert abs(np.sqrt(((f * fl)[1:-1, 1:-1].astype(np.float64) ** 2).sum()) - fnorm) <= 1e-5 * fnorm
    M = port.MG(W, H)
    M.update_fields(fl)
    M.set(p=c["p"], f=f, flag=fl)
    oh = [M.residual(float(O.dx))]
    for _ in range(n):
        M.solve(float(O.dx), zero_gradient_bc=True)
        oh.append(M.residual(float(O.dx)))
    assert np.allclose(hist, oh, rtol=1e-2), (hist, oh)
    # rel_tol <= 0 restores the reference's fixed count
    G.set_tolerance(0.0)
    G.step(dt)
    assert G.solve_info()[0] == 2


def test_graph_replay_equals_direct_launches(ubgl):
    """UBGL_OPT_GRAPH: small grids replay captured CUDA graphs of the fused step after a buffer-role
    state has repeated; fields must equal the directly launched step bit for bit, across sinks,
    a changing dt (part A re-captures, part B keeps replaying) and a flag edit (graphs dropped)."""
    from ubootgl_b200 import capi
    W, H = 258, 131
    c = cases.sim_case(W, H, seed=21)
    dts = [0.001] * 12 + [0.0007] * 8 + [0.001, 0.0007, 0.0009, 0.0009, 0.0009, 0.0009]
    outs, launches = [], []
    for graph in (1, 0):
        s = ubgl.Simulation(c["flag"])
        s.set_option(capi.OPT_GRAPH, graph)
        s.set(capi.VX, c["vx"]); s.set(capi.VY, c["vy"])
        s.set(capi.VX_ACCUM, c["vx_accum"]); s.set(capi.VY_ACCUM, c["vy_accum"])
        for k, dt in enumerate(dts):
            if k == 5:
                s.add_sink(0.4, 0.2, 120.0)
            if k == 15:
                flag2 = c["flag"].copy()
                flag2[40:60, 100:130] = 0
                s.update_flag(flag2)
            s.step(dt)
        outs.append([s.get(f) for f in (capi.VX, capi.VY, capi.P, capi.F, capi.VXB, capi.VYB,
                                        capi.VX_CURRENT, capi.VY_CURRENT, capi.VX_ACCUM, capi.VY_ACCUM)])
        launches.append(s.launch_count())
    for a, b in zip(*outs):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert launches[0] == launches[1]  # replayed launches are counted like direct ones


def test_graph_not_replayed_after_dt_change(ubgl):
    """Round-1 bug (ADVICE): part A's graph was captured with dt = A baked into its kernel arguments;
    after step(dt = B) the same role key (period 3) recurred with dt = B and the stale exec was
    replayed.  Here no flag edit drops the graphs between the two dt phases and every role key
    recurs >= 4 times after each change: graph mode must equal direct launches bit for bit."""
    from ubootgl_b200 import capi
    W, H = 258, 131
    c = cases.sim_case(W, H, seed=23)
    dts = [0.001] * 12 + [0.0006] * 12 + [0.001] * 12
    outs, launches = [], []
    for graph in (1, 0):
        s = ubgl.Simulation(c["flag"])
        s.set_option(capi.OPT_GRAPH, graph)
        s.set(capi.VX, c["vx"]); s.set(capi.VY, c["vy"])
        s.set(capi.VX_ACCUM, c["vx_accum"]); s.set(capi.VY_ACCUM, c["vy_accum"])
        snap = []
        for k, dt in enumerate(dts):
            s.step(dt)
            if k in (11, 14, 17, 23, 26, 35):
                snap.append([s.get(f) for f in (capi.VX, capi.VY, capi.P, capi.F)])
        outs.append(snap)
        launches.append(s.launch_count())
    for sa, sb in zip(*outs):
        for a, b in zip(sa, sb):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert launches[0] == launches[1]


def test_pipelined_host_step_is_the_synchronous_one_a_step_late(ubgl):
    """ubgl_sim_step_host_pipelined: after call n the mirrors hold step n-1, bit for bit what
    ubgl_sim_step_host returned for that step; flush brings the last one; accumulators are consumed
    and cleared per call exactly as in the synchronous form."""
    from ubootgl_b200 import capi
    W, H = 700, 501
    c = cases.sim_case(W, H, seed=31)
    shapes = dict(vx_accum=(H, W - 1), vy_accum=(H - 1, W), vx=(H, W - 1), vy=(H - 1, W), p=(H, W),
                  vx_current=(H, W - 1), vy_current=(H - 1, W))
    sims = []
    for _ in range(2):
        s = ubgl.Simulation(c["flag"])
        s.set(capi.VX, c["vx"]); s.set(capi.VY, c["vy"]); s.set(capi.P, c["p"])
        sims.append(s)
    A, B = sims
    sync_out = []
    steps = 4
    for k in range(steps):
        b = {n: np.full(sh, np.nan, np.float32) for n, sh in shapes.items()}
        b["vx_accum"][:] = c["vx_accum"] * (k + 1)
        b["vy_accum"][:] = c["vy_accum"] * (k + 1)
        A.step_host(0.001, **b)
        sync_out.append(b)
    pb = {n: np.full(sh, np.nan, np.float32) for n, sh in shapes.items()}
    for k in range(steps):
        pb["vx_accum"][:] = c["vx_accum"] * (k + 1)
        pb["vy_accum"][:] = c["vy_accum"] * (k + 1)
        B.step_host_pipelined(0.001, **pb)
        assert not pb["vx_accum"][1:-1, 1:-1].any()  # consumed and cleared
        if k == 0:
            assert np.isnan(pb["vx"]).all()  # nothing to bring down yet
        else:
            for n in ("vx", "vy", "p", "vx_current", "vy_current"):
                assert np.array_equal(pb[n].view(np.uint32), sync_out[k - 1][n].view(np.uint32)), (k, n)
    B.step_host_flush(**{n: pb[n] for n in ("vx", "vy", "p", "vx_current", "vy_current")})
    for n in ("vx", "vy", "p", "vx_current", "vy_current"):
        assert np.array_equal(pb[n].view(np.uint32), sync_out[-1][n].view(np.uint32)), n
    # the device state is the same too
    for f in (capi.VX, capi.VY, capi.P):
        assert np.array_equal(A.get(f).view(np.uint32), B.get(f).view(np.uint32))


def test_current_fields_alias_the_front_until_somebody_writes(ubgl, port):
    """saveCurrentVelocityFields (simulation.cpp:16-19) without the second store: after a fused step
    vx_current / vy_current resolve to the front buffers; a write to either side (upload, stage call,
    setGrids, raw device pointer) first gives *_current its own copy, so the reference's semantics --
    the snapshot keeps the end-of-step values -- hold bit for bit."""
    G, O, c = make_pair(ubgl, port, 130, 97, seed=77)
    dt = float(O.dx)
    for _ in range(2):
        G.step(dt); O.step(dt)
    snap = {f: G.get(f) for f in (ob.VX_CURRENT, ob.VY_CURRENT)}
    for f, g in ((ob.VX_CURRENT, ob.VX), (ob.VY_CURRENT, ob.VY)):
        assert (snap[f].view(np.uint32) == G.get(g).view(np.uint32)).all()
        assert rel_l2(snap[f], O.get(f)) <= 1e-4
    same = lambda f: (G.get(f).view(np.uint32) == snap[f].view(np.uint32)).all()
    # host write to the front buffers: the snapshot stays
    junk = (c["vx"] * 0.5 + 0.25).astype(np.float32)
    G.set(ob.VX, junk); O.set(ob.VX, junk)
    assert same(ob.VX_CURRENT) and same(ob.VY_CURRENT)
    assert (G.get(ob.VX) == junk).all()
    # the next step consumes the written front and makes a new snapshot
    G.step(dt); O.step(dt)
    assert not same(ob.VX_CURRENT)
    snap = {f: G.get(f) for f in (ob.VX_CURRENT, ob.VY_CURRENT)}
    # write to the snapshot itself: the front stays
    front = G.get(ob.VY)
    G.set(ob.VY_CURRENT, np.zeros_like(snap[ob.VY_CURRENT]))
    assert (G.get(ob.VY).view(np.uint32) == front.view(np.uint32)).all()
    assert not G.get(ob.VY_CURRENT).any() and same(ob.VX_CURRENT)
    # a stage call works on the front in place: the snapshot stays
    G.step(dt)
    snap = {f: G.get(f) for f in (ob.VX_CURRENT, ob.VY_CURRENT)}
    G.stage(ob.ST_DIFFUSE, dt)
    assert same(ob.VX_CURRENT) and same(ob.VY_CURRENT)
    assert not (G.get(ob.VX).view(np.uint32) == snap[ob.VX_CURRENT].view(np.uint32)).all()
    # a raw device pointer may be written through: front and snapshot get buffers of their own first
    G.step(dt)
    snap = {f: G.get(f) for f in (ob.VX_CURRENT, ob.VY_CURRENT)}
    pc, _ = G.device_ptr(ob.VX_CURRENT)
    pf, _ = G.device_ptr(ob.VX)
    assert pc != pf and same(ob.VX_CURRENT) and same(ob.VY_CURRENT)
    # setGrids on the device zeroes front velocities in new solids, not the snapshot
    G.step(dt)
    snap = {f: G.get(f) for f in (ob.VX_CURRENT, ob.VY_CURRENT)}
    nf = c["flag"].copy()
    nf[40:60, 30:70] = 0.0
    G.set_grids_all(nf)
    assert same(ob.VX_CURRENT) and same(ob.VY_CURRENT)
