"""Seeded inputs for the SURVEY.md 8f rows (tracers, floating items, terrain
edits), shared by the oracle tests, the GPU parity tests and bench.py."""
import numpy as np

from oracle.bind import ITEM_DTYPE  # dtype only; no oracle code runs here


def make_items(n, W, H, seed, pwidth=0.8, flag=None, cluster=0.0):
    """Debris like explosion.cpp:36-55: size 0.0003*s (s = (1.3..2.3)^2), mass
    0.1*s*(0.1*type+0.15), velocity O(1), spin +-1000.  Positions uniform over
    the interior (a `cluster` fraction is packed into a small patch so that the
    in-bin repulsion has contacts); when `flag` is given, a tenth of the items
    is dropped on solid/fluid borders to exercise the collision branch."""
    rng = np.random.default_rng(seed)
    it = np.zeros(n, ITEM_DTYPE)
    s = (rng.random(n) * 1.0 + 1.3) ** 2
    typ = np.floor(rng.random(n) * 4.0)
    it["size"][:, 0] = 0.0003 * s
    it["size"][:, 1] = 0.0003 * s
    ph = pwidth * H / W
    it["pos"][:, 0] = (0.03 + 0.94 * rng.random(n)) * pwidth
    it["pos"][:, 1] = (0.03 + 0.94 * rng.random(n)) * ph
    nc = int(n * cluster)
    if nc:
        it["pos"][:nc, 0] = 0.31 * pwidth + 0.004 * rng.random(nc)
        it["pos"][:nc, 1] = 0.52 * ph + 0.004 * rng.random(nc)
    if flag is not None:
        # cells whose right neighbour differs: a solid/fluid edge
        ys, xs = np.nonzero(flag[2:-2, 2:-3] != flag[2:-2, 3:-2])
        k = min(n // 10, len(ys))
        if k:
            pick = rng.choice(len(ys), k, replace=False)
            cell = pwidth / W
            it["pos"][-k:, 0] = (xs[pick] + 3.0 + 0.3 * rng.standard_normal(k)) * cell
            it["pos"][-k:, 1] = (ys[pick] + 2.5 + 0.3 * rng.standard_normal(k)) * cell
    it["rotation"] = rng.random(n) * 2 * np.pi
    it["mass"] = 0.1 * s * (0.1 * typ + 0.15)
    ang = rng.random(n) * 2 * np.pi
    it["vel"][:, 0] = 0.5 * rng.standard_normal(n) + np.cos(ang) * 0.1
    it["vel"][:, 1] = 0.5 * rng.standard_normal(n) + np.sin(ang) * 0.1
    it["force"] = (1e-3 * rng.standard_normal((n, 2))).astype(np.float32)
    it["angVel"] = (rng.random(n) - rng.random(n)) * 1000.0
    it["angForce"] = 0.0
    it["bumpCount"] = 0
    return it


def make_bodies(n, W, H, seed, pwidth=0.8, flag=None):
    """Rigid bodies of Simulation::advectFloatingItems (CoItem + CoKinematics): submarines
    (size 0.009 x 0.0022, mass 1.3, ubootgl_app.cpp:94-96) and torpedoes (0.004 x 0.0008, mass
    0.6, launched at player velocity + 0.8, torpedo.cpp:11-17), spin up to +-50
    (ubootgl_app.cpp:40-45), thrust-like forces.  A fifth is dropped next to solid/fluid
    edges (when `flag` is given) to exercise the five terrain probes."""
    rng = np.random.default_rng(seed)
    it = np.zeros(n, ITEM_DTYPE)
    sub = rng.random(n) < 0.5
    it["size"][:, 0] = np.where(sub, 0.009, 0.004)
    it["size"][:, 1] = np.where(sub, 0.0022, 0.0008)
    it["mass"] = np.where(sub, 1.3, 0.6)
    ph = pwidth * H / W
    it["pos"][:, 0] = (0.05 + 0.9 * rng.random(n)) * pwidth
    it["pos"][:, 1] = (0.08 + 0.84 * rng.random(n)) * ph
    if flag is not None:
        ys, xs = np.nonzero(flag[3:-3, 3:-4] != flag[3:-3, 4:-3])
        k = min(n // 5, len(ys))
        if k:
            pick = rng.choice(len(ys), k, replace=False)
            cell = pwidth / W
            it["pos"][-k:, 0] = (xs[pick] + 4.0 + 0.8 * rng.standard_normal(k)) * cell
            it["pos"][-k:, 1] = (ys[pick] + 3.5 + 0.8 * rng.standard_normal(k)) * cell
    it["rotation"] = rng.random(n) * 2 * np.pi
    speed = np.where(sub, 0.3, 0.8) * (0.2 + rng.random(n))
    ang = rng.random(n) * 2 * np.pi
    it["vel"][:, 0] = speed * np.cos(ang)
    it["vel"][:, 1] = speed * np.sin(ang)
    it["force"] = (2.0 * rng.standard_normal((n, 2))).astype(np.float32)
    it["angVel"] = (rng.random(n) - 0.5) * 100.0
    it["angForce"] = (1e-6 * rng.standard_normal(n)).astype(np.float32)
    it["bumpCount"] = 0
    return it


def tracer_state(nt, npts):
    """GLTracers::init (draw_tracers_cs.cpp:29-58): everything zero, ages 2*3.1."""
    return dict(points=np.zeros((nt, npts, 2), np.float32), start=np.zeros(nt, np.uint32),
                end=np.zeros(nt, np.uint32), ages=np.full(nt, 2 * 3.1, np.float32))


def developed_flow(port, W, H, seed, steps=3, dt=None):
    """A few oracle steps from the seeded step-parity case: fields with structure."""
    from oracle import bind as ob
    from tests import cases
    c = cases.sim_case(W, H, seed)
    O = port.Sim(c["flag"])
    O.set(ob.VX, c["vx"])
    O.set(ob.VY, c["vy"])
    for _ in range(steps):
        O.step(dt or 0.001)
    return c["flag"], O
