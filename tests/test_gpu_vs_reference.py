"""GPU (libubgl.so through the C ABI) against the UNMODIFIED reference (oracle/_ref: the
reference's own pressure_solver.cpp / simulation.cpp compiled -Ofast, canonical rbgs path)
DIRECTLY -- no restatement in between -- at BASELINE.json's own sizes:

  configs[1]  the game level, 1090 x 436, from the golden flag bits
  configs[4]  one 4096^2 explosion frame: 16 craters + their sinks + Simulation::step
              (fluid fields and the coarse-flag pyramid)
  configs[2]  one 8192^2 channel step

Tolerances are relative L2 per field, stated beside each assert; measured values are
appended to gpurun_out/parity_errors.jsonl and summarised in profiles/r02_parity_errors.md.
The reference needs OMP threads > H/100 for the canonical red-black order."""
import numpy as np
import pytest

from oracle import bind as ob
from tests import cases, golden_util
from tests.cases import rel_l2

pytestmark = pytest.mark.gpu
PW, MU = 0.8, 0.001
FIELDS = (("vx", ob.VX), ("vy", ob.VY), ("p", ob.P), ("f", ob.F))


def three_way(G, R, Q, what, size):
    """rel-L2 per field of GPU vs reference(-Ofast), GPU vs reference(strict), and the yardstick
    strict vs -Ofast (the reference's own rounding sensitivity); recorded, then asserted:
    the GPU is no further from either build of the reference than max(1e-5, 3 x yardstick)
    (independent rounding perturbations add in quadrature; measured ratios are 1.0-1.5)."""
    e_ofast, e_strict, noise = {}, {}, {}
    for n, f in FIELDS:
        g, r, q = G.get(f), R.get(f), Q.get(f)
        e_ofast[n], e_strict[n], noise[n] = rel_l2(g, r), rel_l2(g, q), rel_l2(q, r)
        del g, r, q
    cases.record_parity(what, size, "unmodified reference (-Ofast)", e_ofast)
    cases.record_parity(what, size, "unmodified reference, strict build", e_strict)
    cases.record_parity(what, size, "yardstick: reference strict vs reference -Ofast", noise)
    for n, _ in FIELDS:
        bound = max(1e-5, 3.0 * noise[n])
        assert e_ofast[n] <= bound, (what, n, e_ofast[n], "noise", noise[n])
        assert e_strict[n] <= bound, (what, n, e_strict[n], "noise", noise[n])
    return e_ofast, e_strict, noise


def pyramid_equal(G, R):
    assert G.mg_levels() == R.mg_levels()
    for l in range(G.mg_levels()):
        a, b = G.mg_flagc(l), R.mg_flagc(l)
        assert a.shape == b.shape and (a.view(np.uint32) == b.view(np.uint32)).all(), l


def test_game_level_steps_vs_unmodified_reference(ubgl, ref, ref_strict):
    flag, pyr, _ = golden_util.game_level()
    H, W = flag.shape
    assert (W, H) == (1090, 436)
    ref.canonical_threads(H)
    ref_strict.canonical_threads(H)
    G, R, Q = ubgl.Simulation(flag, PW, MU), ref.Sim(flag, PW, MU), ref_strict.Sim(flag, PW, MU)
    pyramid_equal(G, R)
    for k in range(3):
        if k == 1:
            for s in (G, R, Q):
                s.add_sink(0.4, 0.15, 120.0)  # explosion.cpp:33
        for s in (G, R, Q):
            s.step(0.001)
        e, _, _ = three_way(G, R, Q, f"game level step {k}", (W, H))
        # at this size the bound of north_star holds without the yardstick for the fields of size O(1)
        assert e["vx"] <= 1e-5 and e["p"] <= 1e-5 and e["f"] <= 1e-5, e


def test_explosion_frame_4096_vs_unmodified_reference(ubgl, ref, ref_strict, port):
    """configs[4] at full size: the fluid side of explosion frames.  The craters are carved on the
    device (ubgl_sim_draw_circles) and, for the reference, into its flag by the restated
    Terrain::drawCircle (bit-exact pinned on terrain.cpp, tests/test_oracle_next.py) followed by
    the reference's own mg.updateFields.  On channel flows the reference's V-cycle AMPLIFIES
    rounding differences (~3x per cycle at this size: its level 0 is all-Neumann and the inflow /
    outflow imbalance makes the problem incompatible), so the yardstick is the
    reference against itself under strict compilation."""
    from bench import CRATERS, crater_list
    S = 4096
    flag, _ = cases.channel_flag(S, S, seed=1234)
    vx, vy = cases.uniform_stream(flag)
    dt = float(np.float32(PW) / np.float32(S - 1))
    h = float(np.float32(PW) / np.float32(S - 1))
    ref.canonical_threads(S)
    ref_strict.canonical_threads(S)
    G, R, Q = ubgl.Simulation(flag, PW, MU), ref.Sim(flag, PW, MU), ref_strict.Sim(flag, PW, MU)
    for s in (G, R, Q):
        s.set(ob.VX, vx)
        s.set(ob.VY, vy)
    g = cases.LCG(99)
    full, simres = flag.copy(), flag.copy()
    for frame in range(2):
        circ = crater_list(g, S, S)
        for cx, cy, d in circ:
            port.draw_circle(full, simres, np.float32(cx), np.float32(cy), int(d), 1.0)
        G.draw_circles(circ, 1.0)
        R.update_flag(simres)
        Q.update_flag(simres)
        for cx, cy, d in circ:
            for s in (G, R, Q):
                s.add_sink(float(np.float32(cx) * np.float32(h)), float(np.float32(cy) * np.float32(h)), 120.0)
        for s in (G, R, Q):
            s.step(dt)
        assert (G.get(ob.FLAG).view(np.uint32) == R.get(ob.FLAG).view(np.uint32)).all()
        pyramid_equal(G, R)
        e, _, _ = three_way(G, R, Q, f"explosion frame {frame} ({CRATERS} craters + sinks + step)", (S, S))
        assert e["f"] <= (1e-5 if frame == 0 else 1.0), e  # the divergence of frame 0 precedes any V-cycle
    assert (simres != flag).sum() > 1000


def test_channel_8192_step_vs_unmodified_reference(ubgl, ref, ref_strict):
    """configs[2], the size BASELINE.json's target is quoted on: one full step, cell by cell."""
    S = 8192
    flag, _ = cases.channel_flag(S, S, seed=1234)
    vx, vy = cases.uniform_stream(flag)
    dt = float(np.float32(PW) / np.float32(S - 1))
    ref.canonical_threads(S)  # 82 threads: every level takes the canonical red-black order
    ref_strict.canonical_threads(S)
    G, R, Q = ubgl.Simulation(flag, PW, MU), ref.Sim(flag, PW, MU), ref_strict.Sim(flag, PW, MU)
    for s in (G, R, Q):
        s.set(ob.VX, vx)
        s.set(ob.VY, vy)
    del vx, vy
    for s in (G, R, Q):
        s.step(dt)
    pyramid_equal(G, R)
    e, _, _ = three_way(G, R, Q, "channel step 0", (S, S))
    assert e["f"] <= 1e-5, e  # advect + divergence, before the V-cycles
