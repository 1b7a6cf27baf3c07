"""GPU: the fused / temporally blocked kernels must reproduce the plain
one-kernel-per-reference-stage path BIT FOR BIT (same per-cell arithmetic from
stencils.cuh, only the schedule differs)."""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def bits(a):
    return a.view(np.uint32)


def same(a, b):
    # +0 / -0 are both "zero" for every consumer; everything else must match bitwise
    return bool(((bits(a) == bits(b)) | ((a == 0) & (b == 0))).all())


MG_SIZES = [(16, 16), (17, 23), (40, 33), (112, 48), (113, 49), (128, 64), (131, 77), (224, 96),
            (225, 97), (257, 130), (300, 301), (545, 218), (1090, 436), (1025, 1025)]


@pytest.mark.parametrize("W,H", MG_SIZES)
@pytest.mark.parametrize("zg", [True, False])
def test_vcycle_fused_equals_plain(ubgl, W, H, zg):
    flag, g = cases.channel_flag(W, H, seed=W + 3 * H, ndiscs=10, radius=max(2.0, H / 14.0),
                                 closed_box=(W % 2 == 0))
    rng = np.random.default_rng(W * H)
    f = rng.standard_normal((H, W)).astype(np.float32)
    p0 = rng.standard_normal((H, W)).astype(np.float32)
    hh = np.float32(0.8 / (W - 1))
    out = []
    for fused in (1, 2, 0):  # shared-memory tile kernels, register-run kernels (left selected), plain
        m = ubgl.MG(W, H)
        m.set_option(ubgl.capi.OPT_FUSED, fused)
        m.update_fields(flag)
        m.set(p0, f, flag)
        m.solve(hh, zg, 2)
        out.append(m.get_p())
    assert same(out[1], out[2]), ("run vs plain", np.abs(out[1] - out[2]).max())
    assert same(out[0], out[2]), ("tile vs plain", np.abs(out[0] - out[2]).max())


def test_nonbinary_flags_fall_back_to_plain(ubgl, port):
    """A flag field that is not {0,1} cannot use the bit-mask kernels; the
    library must notice and still match the oracle."""
    W, H = 131, 77
    rng = np.random.default_rng(1)
    flag = (0.25 + 0.75 * rng.random((H, W))).astype(np.float32)
    f = rng.standard_normal((H, W)).astype(np.float32)
    p0 = np.zeros((H, W), np.float32)
    m = ubgl.MG(W, H)
    m.update_fields(flag)
    m.set(p0, f, flag)
    m.solve(0.01, True, 1)
    o = port.MG(W, H)
    o.update_fields(flag)
    o.set(p0, f, flag)
    o.solve(0.01, True)
    assert cases.rel_l2(m.get_p(), o.get_p()) <= 1e-5


@pytest.mark.parametrize("W,H", [(70, 40), (130, 97), (258, 131), (545, 218), (1090, 436), (700, 501), (1301, 300), (2048, 1536)])
def test_step_fused_equals_plain(ubgl, W, H):
    """(2048, 1536) is large enough for the register-run prestep strips of 16 rows (k_prestep_run<., 16>);
    the 64-row strips run in tests/test_gpu_fullsize.py at 8192^2."""
    from ubootgl_b200 import capi
    c = cases.sim_case(W, H, seed=W + H)
    outs = []
    for fused in (1, 2, 0):
        s = ubgl.Simulation(c["flag"])
        s.set_option(capi.OPT_FUSED, fused)
        s.set(capi.VX, c["vx"]); s.set(capi.VY, c["vy"])
        s.set(capi.VX_ACCUM, c["vx_accum"]); s.set(capi.VY_ACCUM, c["vy_accum"])
        s.add_sink(0.4, 0.4 * H / W, 120.0)
        for _ in range(3):
            s.step(0.001)
        outs.append([s.get(f) for f in (capi.VX, capi.VY, capi.P, capi.F, capi.VXB, capi.VYB,
                                        capi.VX_CURRENT, capi.VY_CURRENT, capi.VX_ACCUM)])
    for a, b, c_ in zip(*outs):
        assert same(b, c_), ("run vs plain", np.abs(b - c_).max())
        assert same(a, c_), ("tile vs plain", np.abs(a - c_).max())
