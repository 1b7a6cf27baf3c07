"""GPU parity of the callers either side of the step: tracers +
co-located velocity, floating items with accumulator scatter, terrain edits and
scrolling -- libubgl.so through the C ABI against the CPU oracle
(oracle/ubgl_oracle_next.c) on identical seeded inputs."""
import numpy as np
import pytest

from oracle import bind as ob
from tests import cases, next_cases
from tests.cases import rel_l2

pytestmark = pytest.mark.gpu


def gpu_twin(ubgl, O, flag):
    """A device Simulation holding exactly the oracle simulation's fields."""
    G = ubgl.Simulation(flag)
    for f in (ob.VX, ob.VY, ob.VXB, ob.VYB, ob.P, ob.VX_ACCUM, ob.VY_ACCUM, ob.VX_CURRENT, ob.VY_CURRENT):
        G.set(f, O.get(f))
    return G


@pytest.mark.parametrize("W,H", [(70, 40), (130, 97), (258, 131)])
def test_colocate_bit_exact(ubgl, port, W, H):
    flag, O = next_cases.developed_flow(port, W, H, seed=W)
    G = gpu_twin(ubgl, O, flag)
    vxy, mag = G.colocate_velocity()
    ovxy, omag = port.colocate(O.get(ob.VX_CURRENT), O.get(ob.VY_CURRENT))
    assert (vxy.view(np.uint32) == ovxy.view(np.uint32)).all()
    assert (mag.view(np.uint32) == omag.view(np.uint32)).all()


@pytest.mark.parametrize("W,H,nt", [(130, 97, 2000), (1090, 436, 20000)])
def test_tracers_and_texture_against_the_shader_source(ubgl, port, glsl, W, H, nt):
    """The GPU directly against the reference's GLSL compute shaders (interp_shader.cs,
    advect_tracer_points.cs: unmodified sources compiled through oracle/shim/glsl_shim.hpp): the co-located
    velocity / magnitude texture and 60 tracer frames, bit for bit."""
    flag, O = next_cases.developed_flow(port, W, H, seed=W + 7)
    G = gpu_twin(ubgl, O, flag)
    vxy, mag = G.colocate_velocity()
    svxy, smag = glsl.colocate(O.get(ob.VX_CURRENT), O.get(ob.VY_CURRENT))
    assert (vxy.view(np.uint32) == svxy.view(np.uint32)).all()
    assert (mag.view(np.uint32) == smag.view(np.uint32)).all()
    T = ubgl.Tracers(nt, 30)
    st = next_cases.tracer_state(nt, 30)
    pd = (np.float32(0.8), np.float32(0.8) * np.float32(H) / np.float32(W))
    g = cases.LCG(3)
    for k in range(60):
        seed = int(g.u() * 2 ** 31)
        T.advect(G, 0.02, seed)
        glsl.tracers_advect(st, 0.02, pd, seed, svxy, flag)
    gs = T.state()
    for key in ("points", "ages"):
        assert (gs[key].view(np.uint32) == st[key].view(np.uint32)).all(), key
    for key in ("start", "end"):
        assert (gs[key] == st[key]).all(), key


@pytest.mark.parametrize("W,H", [(70, 40), (258, 131), (1090, 436)])
def test_display_export_into_cuda_arrays(ubgl, port, W, H):
    """8f rank 4: the texels the reference uploads per frame (velocity_textures.cpp:63-93 through
    interp_shader.cs, p through draw_2dbuf.cpp:181-209) written into CUDA arrays -- the form a mapped
    GL texture takes -- without leaving the device: bit-identical to the restated shader and to p."""
    flag, O = next_cases.developed_flow(port, W, H, seed=W + 1)
    G = gpu_twin(ubgl, O, flag)
    tw, th = 2 * W - 1, 2 * H - 1
    avxy, amag, ap = ubgl.DisplayArray(tw, th, 2), ubgl.DisplayArray(tw, th, 1), ubgl.DisplayArray(W, H, 1)
    G.export_display(avxy, amag, ap)
    ovxy, omag = port.colocate(O.get(ob.VX_CURRENT), O.get(ob.VY_CURRENT))
    assert (avxy.read().view(np.uint32) == ovxy.view(np.uint32)).all()
    assert (amag.read().view(np.uint32) == omag.view(np.uint32)).all()
    assert (ap.read().view(np.uint32) == O.get(ob.P).view(np.uint32)).all()
    # any subset; after a step the arrays follow the new fields
    p_before = ap.read()
    G.step(0.001)
    G.export_display(None, amag, None)
    _, omag2 = port.colocate(G.get(ob.VX_CURRENT), G.get(ob.VY_CURRENT))
    assert (amag.read().view(np.uint32) == omag2.view(np.uint32)).all()
    assert (ap.read().view(np.uint32) == p_before.view(np.uint32)).all()  # p was not exported this time
    # wrong texel layout or size: rejected, nothing written
    with pytest.raises(ubgl.UbglError):
        G.export_display(amag, None, None)  # R32F where RG32F is expected
    with pytest.raises(ubgl.UbglError):
        G.export_display(None, None, amag)  # (2W-1) x (2H-1) where W x H is expected
    for a in (avxy, amag, ap):
        a.close()


@pytest.mark.parametrize("W,H,nt,scale", [(130, 97, 1000, 1), (258, 131, 5000, 1), (130, 97, 2000, 2)])
def test_tracers_bit_exact(ubgl, port, W, H, nt, scale):
    """60 frames incl. respawn, ring wrap-around, freezing in terrain: bit-exact
    (the on-the-fly texel evaluation equals sampling the materialised texture)."""
    flag, O = next_cases.developed_flow(port, W, H, seed=H)
    G = gpu_twin(ubgl, O, flag)
    flagtex = np.kron(flag, np.ones((scale, scale), np.float32))  # terrain.flagFullRes at `scale`
    ovxy, _ = port.colocate(O.get(ob.VX_CURRENT), O.get(ob.VY_CURRENT))
    T = ubgl.Tracers(nt, 30)
    if scale != 1:
        T.set_flag_texture(flagtex)
    st = next_cases.tracer_state(nt, 30)
    pd = (np.float32(0.8), np.float32(0.8) * np.float32(H) / np.float32(W))
    g = cases.LCG(5)
    for k in range(60):
        seed = int(g.u() * 2 ** 31)
        T.advect(G, 0.02, seed)
        port.tracers_advect(st, 0.02, pd, seed, ovxy, flagtex)
    gs = T.state()
    for key in ("points", "ages"):
        assert (gs[key].view(np.uint32) == st[key].view(np.uint32)).all(), key
    assert (gs["start"] == st["start"]).all() and (gs["end"] == st["end"]).all()
    assert (st["ages"] > 1.0).any() and (st["end"] > 20).any()
    T.shift(-0.125)
    port.tracers_shift(st, -0.125)
    assert (T.state()["points"] == st["points"]).all()


@pytest.mark.parametrize("W,H,n,dt", [(130, 97, 500, 0.004), (258, 131, 4000, 0.01), (70, 40, 64, 0.02)])
def test_items_match_oracle(ubgl, port, W, H, n, dt):
    flag, O = next_cases.developed_flow(port, W, H, seed=W + H)
    G = gpu_twin(ubgl, O, flag)
    items = next_cases.make_items(n, W, H, seed=3, flag=flag, cluster=0.2)
    I = ubgl.Items(items)
    o = items.copy()
    ax, ay = O.get(ob.VX_ACCUM), O.get(ob.VY_ACCUM)
    vx, vy, p = O.get(ob.VX), O.get(ob.VY), O.get(ob.P)
    for k in range(3):
        I.advect_simple(G, dt)
        port.items_advect_simple(o, dt, flag, vx, vy, p, ax, ay)
        g = I.get()
        for name in ("pos", "vel", "rotation", "angVel", "force", "angForce", "size", "mass"):
            assert rel_l2(g[name], o[name]) <= 1e-5, (k, name, rel_l2(g[name], o[name]))
        assert (g["bumpCount"] == o["bumpCount"]).mean() >= 0.995
    assert o["bumpCount"].sum() > 0
    # force scatter: device atomicAdd vs the oracle's serial +=
    assert np.abs(ax).sum() > 0
    assert rel_l2(G.get(ob.VX_ACCUM), ax) <= 1e-5 and rel_l2(G.get(ob.VY_ACCUM), ay) <= 1e-5
    # and the accumulators feed the next step exactly like host-uploaded ones
    O.set(ob.VX_ACCUM, ax); O.set(ob.VY_ACCUM, ay)
    G.step(0.001); O.step(0.001)
    for f in (ob.VX, ob.VY, ob.P):
        assert rel_l2(G.get(f), O.get(f)) <= 3e-5


# share of rigid bodies that must agree with the oracle record for record (the rest sit within
# rounding of a terrain-probe threshold and take the other branch); see profiles/r02_parity_errors.md
BODY_CLOSE = 0.999  # measured 1.0 in every frame since the rotation is correctly rounded (round 1: >= 0.99)


@pytest.mark.parametrize("W,H,n,dt", [(130, 97, 300, 0.004), (258, 131, 1000, 0.01), (545, 218, 400, 1.0 / 60.0)])
def test_bodies_match_oracle(ubgl, port, W, H, n, dt):
    """ubgl_items_advect (Simulation::advectFloatingItems, rigid bodies) against the restatement
    that tests/test_oracle_next.py pins on the unmodified reference."""
    from tests.test_oracle_next import close_fraction
    flag, O = next_cases.developed_flow(port, W, H, seed=W + H)
    G = gpu_twin(ubgl, O, flag)
    items = next_cases.make_bodies(n, W, H, seed=5, flag=flag)
    I = ubgl.Items(items)
    o = items.copy()
    ax, ay = O.get(ob.VX_ACCUM), O.get(ob.VY_ACCUM)
    vx, vy = O.get(ob.VX), O.get(ob.VY)
    for k in range(3):
        I.advect(G, dt)
        port.items_advect(o, dt, flag, vx, vy, ax, ay)
        g = I.get()
        assert np.isfinite(g["pos"]).all()
        cf, bf = close_fraction(g, o), float((g["bumpCount"] == o["bumpCount"]).mean())
        cases.record_parity(f"rigid bodies frame {k}: share of {n} bodies within 2e-4 / equal bumpCount", (W, H),
                            "restatement (= unmodified reference to 1e-7)", {"close": cf, "bump": bf})
        assert cf >= BODY_CLOSE, (k, cf)
        assert bf >= BODY_CLOSE
        assert (g["force"] == 0).all() and (g["angForce"] == 0).all()
    assert o["bumpCount"].sum() > 0
    assert np.abs(ax).sum() > 0
    eax, eay = rel_l2(G.get(ob.VX_ACCUM), ax), rel_l2(G.get(ob.VY_ACCUM), ay)
    # the scattered reaction feeds the next fluid step like host-uploaded accumulators
    O.set(ob.VX_ACCUM, ax); O.set(ob.VY_ACCUM, ay)
    G.step(0.001); O.step(0.001)
    es = {n: rel_l2(G.get(f), O.get(f)) for n, f in (("vx", ob.VX), ("vy", ob.VY), ("p", ob.P))}
    cases.record_parity(f"rigid bodies: accumulators after 3 frames, fields after the next step ({n} bodies)", (W, H),
                        "restatement (= unmodified reference to 1e-7)", dict(vx_accum=eax, vy_accum=eay, **es))
    assert eax <= 1e-4 and eay <= 1e-4, (eax, eay)  # atomicAdd order vs the serial +=
    for k_, e in es.items():
        assert e <= 1e-4, (k_, e)


def test_items_empty_and_regrow(ubgl, port):
    flag, O = next_cases.developed_flow(port, 70, 40, seed=1)
    G = gpu_twin(ubgl, O, flag)
    I = ubgl.Items()
    I.set(np.zeros(0, ubgl.capi.ITEM_DTYPE))
    I.advect_simple(G, 0.01)
    assert len(I.get()) == 0
    I.set(next_cases.make_items(10, 70, 40, seed=1))
    I.advect_simple(G, 0.01)
    I.set(next_cases.make_items(300, 70, 40, seed=2))
    I.advect_simple(G, 0.01)
    assert np.isfinite(I.get()["pos"]).all()


def test_draw_circles_bit_exact_and_pyramid(ubgl, port):
    W, H = 258, 131
    flag, O = next_cases.developed_flow(port, W, H, seed=9)
    G = gpu_twin(ubgl, O, flag)
    g = cases.LCG(99)
    full, simres = flag.copy(), flag.copy()
    for val in (1.0, 0.0, 1.0):
        circ = []
        for _ in range(16):
            d = 2 + int(g.u() * 10)
            circ.append((d + 1 + g.u() * (W - 2 * d - 3), d + 1 + g.u() * (H - 2 * d - 3), d))
        for cx, cy, d in circ:
            port.draw_circle(full, simres, np.float32(cx), np.float32(cy), d, val)
        G.draw_circles(circ, val)
        assert (G.get(ob.FLAG) == simres).all()
        O.update_flag(simres)
        for l in range(G.mg_levels()):
            assert (G.mg_flagc(l) == O.mg_flagc(l)).all(), l
    assert (simres != flag).sum() > 200
    G.step(0.001); O.step(0.001)
    for f in (ob.VX, ob.VY, ob.P):
        assert rel_l2(G.get(f), O.get(f)) <= 3e-5
    with pytest.raises(ubgl.UbglError):
        G.draw_circles([(3.0, 50.0, 10)], 1.0)  # box leaves the grid


@pytest.mark.parametrize("W,H", [(70, 40), (258, 131), (1090, 436)])
def test_shift_map_bit_exact(ubgl, port, W, H):
    flag, O = next_cases.developed_flow(port, W, H, seed=W, steps=2)
    G = gpu_twin(ubgl, O, flag)
    rng = np.random.default_rng(W)
    f = {k: O.get(i) for k, i in dict(vxf=ob.VX, vxb=ob.VXB, vyf=ob.VY, vyb=ob.VYB, p=ob.P,
                                      vxc=ob.VX_CURRENT, vyc=ob.VY_CURRENT).items()}
    cur = flag.copy()
    for k in range(3):
        col = (rng.random(H) > 0.4).astype(np.float32)
        new = np.roll(cur, -1, axis=1)
        new[:, -1] = col
        port.shift_map(cur, f["vxf"], f["vxb"], f["vyf"], f["vyb"], f["p"], f["vxc"], f["vyc"], new)
        G.shift_map(col)
        assert (cur == new).all()
    assert (G.get(ob.FLAG) == cur).all()
    for k, i in dict(vxf=ob.VX, vxb=ob.VXB, vyf=ob.VY, vyb=ob.VYB, p=ob.P, vxc=ob.VX_CURRENT,
                     vyc=ob.VY_CURRENT).items():
        assert (G.get(i).view(np.uint32) == f[k].view(np.uint32)).all(), k
    O.update_flag(cur)
    for l in range(G.mg_levels()):
        assert (G.mg_flagc(l) == O.mg_flagc(l)).all()


def test_set_grids_all(ubgl, port):
    W, H = 130, 97
    flag, O = next_cases.developed_flow(port, W, H, seed=4)
    G = gpu_twin(ubgl, O, flag)
    new = flag.copy()
    new[30:50, 40:70] = 0
    new[60:70, 20:30] = 1
    fl, vx, vy, p = flag.copy(), O.get(ob.VX), O.get(ob.VY), O.get(ob.P)
    port.set_grids_all(fl, vx, vy, p, new)
    G.set_grids_all(new)
    assert (G.get(ob.FLAG) == new).all()
    assert (G.get(ob.VX) == vx).all() and (G.get(ob.VY) == vy).all() and (G.get(ob.P) == p).all()
