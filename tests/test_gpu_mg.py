"""GPU parity: multigrid operators and V-cycle (libubgl.so through the C ABI)
against the CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): fp32 relative-L2 <= 1e-5 per stage,
V-cycle residual history within 1 %, flag masks bit-exact."""
import numpy as np
import pytest

from tests import cases
from tests.cases import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-5
SIZES = [(8, 8), (9, 17), (33, 20), (64, 64), (131, 77), (257, 130), (300, 301)]


@pytest.mark.parametrize("W,H", SIZES)
def test_rbgs(ubgl, port, W, H):
    flag, p, f = cases.random_fields(W, H, seed=W * 1000 + H)
    for sweeps in (1, 3):
        got = ubgl.capi.rbgs(p, f, flag, 0.01, 1.0, sweeps)
        want = port.rbgs(p, f, flag, 0.01, 1.0, sweeps)
        assert rel_l2(got, want) <= TOL
        # boundary ring is never written by rbgs
        assert (got[0] == p[0]).all() and (got[:, 0] == p[:, 0]).all()
        assert (got[-1] == p[-1]).all() and (got[:, -1] == p[:, -1]).all()


def test_rbgs_alpha(ubgl, port):
    flag, p, f = cases.random_fields(65, 40, seed=5)
    got = ubgl.capi.rbgs(p, f, flag, 0.02, 0.7, 2)
    want = port.rbgs(p, f, flag, 0.02, 0.7, 2)
    assert rel_l2(got, want) <= TOL


@pytest.mark.parametrize("W,H", SIZES)
def test_residual_restrict(ubgl, port, W, H):
    flag, p, f = cases.random_fields(W, H, seed=W * 7 + H)
    r_g, l2_g = ubgl.capi.residual(p, f, flag, 0.01)
    r_o, l2_o = port.residual(p, f, flag, 0.01)
    assert rel_l2(r_g, r_o) <= TOL
    assert abs(l2_g - l2_o) <= 1e-5 * l2_o
    rc_g = ubgl.capi.restrict(r_o)
    rc_o = port.restrict(r_o)
    assert rc_g.shape == rc_o.shape
    assert rel_l2(rc_g, rc_o) <= TOL
    assert (rc_g[0] == 0).all() and (rc_g[:, 0] == 0).all()


@pytest.mark.parametrize("W,H", SIZES)
def test_prolongate_correct_bc(ubgl, port, W, H):
    rng = np.random.default_rng(W + 31 * H)
    flag, p, _ = cases.random_fields(W, H, seed=W + H)
    flagc = (rng.random((H // 2, W // 2)) > 0.3).astype(np.float32)
    ec = rng.standard_normal((H // 2, W // 2)).astype(np.float32)
    e_g = ubgl.capi.prolongate(ec, flagc, flag)
    e_o = port.prolongate(ec, flagc, flag)
    assert rel_l2(e_g, e_o) <= TOL
    assert ((e_g == 0) == (e_o == 0)).all()  # same untouched cells
    assert rel_l2(ubgl.capi.correct(p, e_o), port.correct(p, e_o)) <= TOL
    assert (ubgl.capi.zero_gradient_bc(p) == port.zero_gradient_bc(p)).all()


@pytest.mark.parametrize("W,H", [(8, 8), (16, 9), (131, 77), (545, 218), (1090, 436), (1025, 1025)])
def test_flag_pyramid_bit_exact(ubgl, port, W, H):
    flag, _ = cases.channel_flag(W, H, seed=W + H, ndiscs=12, radius=max(2.0, H / 12.0))
    rng = np.random.default_rng(W)
    flag[(rng.random((H, W)) > 0.97)] = 0  # speckle to exercise the 0.2 threshold
    g = ubgl.MG(W, H)
    o = port.MG(W, H)
    g.update_fields(flag)
    o.update_fields(flag)
    assert g.levels() == o.levels() == len(cases_levels(W, H))
    for l in range(g.levels()):
        a, b = g.flagc(l), o.flagc(l)
        assert a.shape == b.shape
        assert (a.view(np.uint32) == b.view(np.uint32)).all(), f"level {l}"


def cases_levels(W, H):
    from oracle import bind
    return bind.mg_level_sizes(W, H)


@pytest.mark.parametrize("W,H,zg", [(64, 64, True), (131, 77, True), (131, 77, False),
                                    (257, 130, True), (545, 218, True), (1090, 436, True)])
def test_vcycle_vs_oracle(ubgl, port, W, H, zg):
    """Three V-cycles on a channel-with-obstacles problem: p within 1e-5 after
    the first cycle, residual history within 1 %."""
    flag, g = cases.channel_flag(W, H, seed=7, ndiscs=8, radius=H / 12.0, closed_box=True)
    f = cases.dipole_rhs(flag, g, n=32)
    hh = np.float32(0.8 / (W - 1))
    p0 = np.zeros((H, W), np.float32)
    G = ubgl.MG(W, H)
    O = port.MG(W, H)
    G.update_fields(flag)
    O.update_fields(flag)
    G.set(p0, f, flag)
    O.set(p0, f, flag)
    r0 = O.residual(hh)
    assert abs(G.residual(hh) - r0) <= 1e-5 * r0
    for cyc in range(3):
        G.solve(hh, zg, 1)
        O.solve(hh, zg)
        assert rel_l2(G.get_p(), O.get_p()) <= TOL * (1 + cyc)
        rg, ro = G.residual(hh), O.residual(hh)
        assert abs(rg / r0 - ro / r0) <= 0.01 * (ro / r0), (cyc, rg, ro)


def test_mg_solve_host_matches_resident(ubgl):
    W, H = 131, 77
    flag, g = cases.channel_flag(W, H, seed=3, ndiscs=5, radius=6.0, closed_box=True)
    f = cases.dipole_rhs(flag, g, n=16)
    hh = np.float32(0.8 / (W - 1))
    A = ubgl.MG(W, H)
    A.update_fields(flag)
    p = np.zeros((H, W), np.float32)
    A.solve_host(p, f, flag, hh, True)
    B = ubgl.MG(W, H)
    B.update_fields(flag)
    B.set(np.zeros((H, W), np.float32), f, flag)
    B.solve(hh, True, 1)
    assert (p == B.get_p()).all()


def test_mgtest_known_answer(ubgl):
    """mgtest.cpp:10-53 on the GPU against the reference's own canonical-path
    history (survey KAT, reproduced by oracle/_ref in tests/golden)."""
    import json, os
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mgtest_kat.json")))
    u, rhs, flag, h, ref = cases.mgtest_problem(1025)
    G = ubgl.MG(1025, 1025)
    G.set(u, rhs, flag)
    hist = [G.residual(h)]
    for _ in range(5):
        G.solve(h, False, 1)
        hist.append(G.residual(h))
    for a, b in zip(hist, kat["residual_history"]):
        assert abs(a - b) <= 0.01 * b, (hist, kat["residual_history"])
    err = cases.mgtest_error(ref, G.get_p())
    assert abs(err - kat["scaled_error"]) <= 0.01 * kat["scaled_error"]


def test_errors_are_loud(ubgl):
    with pytest.raises(ubgl.UbglError):
        ubgl.MG(4, 4)
    with pytest.raises(ubgl.UbglError):
        ubgl.Simulation(np.ones((5, 5), np.float32))
    s = ubgl.Simulation(np.ones((16, 16), np.float32))
    with pytest.raises(ubgl.UbglError):
        s.set(ubgl.capi.VX, np.zeros((16, 16), np.float32))
