"""Seeded synthetic inputs shared by the parity tests, the golden-fixture
generator and bench.py.  numpy only."""
import numpy as np


class LCG:
    """s = 1664525*s + 1013904223 (mod 2^32), u = (s >> 8) / 2^24."""

    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFF

    def u(self):
        self.s = (1664525 * self.s + 1013904223) & 0xFFFFFFFF
        return (self.s >> 8) / float(1 << 24)


def channel_flag(W, H, seed=1234, ndiscs=32, radius=None, closed_box=False):
    """Channel with obstacles: rows 0 and H-1 solid, `ndiscs` discs of radius
    H/32 at LCG centres ((0.1+0.8u)W, (0.15+0.7u)H), clipped to the interior."""
    flag = np.ones((H, W), np.float32)
    flag[0, :] = 0
    flag[H - 1, :] = 0
    if closed_box:
        flag[:, 0] = 0
        flag[:, W - 1] = 0
    g = LCG(seed)
    r = (H / 32.0) if radius is None else radius
    for _ in range(ndiscs):
        cx = (0.1 + 0.8 * g.u()) * W
        cy = (0.15 + 0.7 * g.u()) * H
        x0, x1 = max(1, int(np.floor(cx - r))), min(W - 2, int(np.floor(cx + r)))
        y0, y1 = max(1, int(np.floor(cy - r))), min(H - 2, int(np.floor(cy + r)))
        if x1 < x0 or y1 < y0:
            continue
        yy, xx = np.mgrid[y0:y1 + 1, x0:x1 + 1]
        m = (xx - cx) ** 2 + (yy - cy) ** 2 <= r * r
        sub = flag[y0:y1 + 1, x0:x1 + 1]
        sub[m] = 0
    return flag, g


def channel_flag_rows(W, H, y0, y1, seed=1234, ndiscs=32, radius=None, closed_box=False):
    """Rows [y0, y1) of channel_flag(W, H, ...) without building the whole grid
    (the 32768^2 slab runs generate only the rows a rank stores)."""
    rows = np.ones((y1 - y0, W), np.float32)
    if y0 <= 0 < y1:
        rows[0 - y0, :] = 0
    if y0 <= H - 1 < y1:
        rows[H - 1 - y0, :] = 0
    if closed_box:
        rows[:, 0] = 0
        rows[:, W - 1] = 0
    g = LCG(seed)
    r = (H / 32.0) if radius is None else radius
    for _ in range(ndiscs):
        cx = (0.1 + 0.8 * g.u()) * W
        cy = (0.15 + 0.7 * g.u()) * H
        x0, x1 = max(1, int(np.floor(cx - r))), min(W - 2, int(np.floor(cx + r)))
        ya, yb = max(1, int(np.floor(cy - r))), min(H - 2, int(np.floor(cy + r)))
        ya, yb = max(ya, y0), min(yb, y1 - 1)
        if x1 < x0 or yb < ya:
            continue
        yy, xx = np.mgrid[ya:yb + 1, x0:x1 + 1]
        m = (xx - cx) ** 2 + (yy - cy) ** 2 <= r * r
        sub = rows[ya - y0:yb + 1 - y0, x0:x1 + 1]
        sub[m] = 0
    return rows


def channel_row_fluid_fraction(W, H, seed=1234, ndiscs=32, radius=None):
    """Fluid fraction of every row of channel_flag(W, H, ...) from the disc list alone (chords;
    overlapping discs are counted twice, which is immaterial for load balancing)."""
    g = LCG(seed)
    r = (H / 32.0) if radius is None else radius
    solid = np.zeros(H, np.float64)
    y = np.arange(H, dtype=np.float64)
    for _ in range(ndiscs):
        cx = (0.1 + 0.8 * g.u()) * W
        cy = (0.15 + 0.7 * g.u()) * H
        d2 = r * r - (y - cy) ** 2
        solid += np.where(d2 > 0, 2.0 * np.sqrt(np.maximum(d2, 0.0)), 0.0)
    frac = 1.0 - np.minimum(solid, W) / W
    frac[0] = frac[H - 1] = 0.0
    return frac


def dipole_rhs(flag, g, n=64, amp=1000.0):
    """Zero-sum dipoles f(x,y)=+amp, f(x+3,y)=-amp at LCG positions in fluid."""
    H, W = flag.shape
    f = np.zeros((H, W), np.float32)
    for _ in range(n):
        x = W // 8 + int(g.u() * (3 * W // 4))
        y = H // 8 + int(g.u() * (3 * H // 4))
        if x + 3 < W and flag[y, x] == 1 and flag[y, x + 3] == 1:
            f[y, x] = amp
            f[y, x + 3] = -amp
    return f


def uniform_stream(flag):
    """vx(x,y) = flag(x,y)*flag(x+1,y), inflow column vx(0,y)=1; vy = 0."""
    H, W = flag.shape
    vx = (flag[:, :-1] * flag[:, 1:]).astype(np.float32)
    vx[:, 0] = 1.0
    vy = np.zeros((H - 1, W), np.float32)
    return vx, vy


def random_fields(W, H, seed, solid=0.25):
    rng = np.random.default_rng(seed)
    flag = (rng.random((H, W)) > solid).astype(np.float32)
    p = rng.standard_normal((H, W)).astype(np.float32)
    f = rng.standard_normal((H, W)).astype(np.float32)
    return flag, p, f


def sim_case(W, H, seed, ndiscs=6, radius=None):
    """Flag + perturbed velocity + accumulators for step-level parity."""
    flag, _ = channel_flag(W, H, seed=seed, ndiscs=ndiscs, radius=radius or H / 10.0)
    rng = np.random.default_rng(seed)
    vx = (1.0 + 0.3 * rng.standard_normal((H, W - 1))).astype(np.float32)
    vy = (0.3 * rng.standard_normal((H - 1, W))).astype(np.float32)
    ax = (0.01 * rng.standard_normal((H, W - 1))).astype(np.float32)
    ay = (0.01 * rng.standard_normal((H - 1, W))).astype(np.float32)
    p = (0.1 * rng.standard_normal((H, W))).astype(np.float32)
    return dict(flag=flag, vx=vx, vy=vy, vx_accum=ax, vy_accum=ay, p=p)


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-30))


def mgtest_problem(N=1025):
    """mgtest.cpp:10-32: Laplace, Dirichlet sinh(pi)sin(pi x) on the top row."""
    import math
    h = np.float32(1.0 / (N - 1))
    u = np.zeros((N, N), np.float32)
    x = np.arange(1, N - 1)
    u[N - 1, 1:N - 1] = (math.sinh(math.pi) * np.sin(x / (N - 1.0) * math.pi)).astype(np.float32)
    rhs = np.zeros((N, N), np.float32)
    flag = np.ones((N, N), np.float32)
    yy = np.arange(N)[:, None] * float(h) * math.pi
    xx = np.arange(N)[None, :] * float(h) * math.pi
    ref = (np.sinh(yy) * np.sin(xx)).astype(np.float32)
    ref[:, 0] = 0
    ref[:, N - 1] = 0
    return u, rhs, flag, h, ref


def mgtest_error(ref, u):
    """mgtest.cpp:44-53: sqrt(sum err^2)/N/N."""
    N = u.shape[0]
    err = (ref - u).astype(np.float64)
    return float(np.sqrt((err ** 2).sum())) / N / N


def record_parity(what, size, against, errors):
    """Append one line of measured parity errors to gpurun_out/parity_errors.jsonl (summarised
    under profiles/ per round).  Best effort: the tests never depend on it."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_errors.jsonl"), "a") as fp:
            fp.write(json.dumps({"what": what, "size": list(size), "against": against,
                                 "rel_l2": {k: float(v) for k, v in errors.items()}}) + "\n")
    except OSError:
        pass
