"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the two CPU checkers.

* ``Ref``  : oracle/_ref/libubgl_ref.so, the UNMODIFIED reference translation
             units (pressure_solver.cpp, simulation.cpp, terrain.cpp) compiled
             where they lie under /root/reference by oracle/Makefile, behind
             the forwarding wrapper oracle/ref_capi.cpp.
* ``Port`` : oracle/_build/liboracle.so, the plain-C restatement
             oracle/ubgl_oracle.c (each function cites the reference lines it
             follows).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product path
(ubootgl_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libubgl_ref.so")
REF_STRICT_SO = os.path.join(HERE, "_ref", "libubgl_ref_strict.so")
PORT_SO = os.path.join(HERE, "_build", "liboracle.so")

FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int)

# field / stage ids shared by ref_capi.cpp, ubgl_oracle.h and include/ubgl.h
FLAG, VX, VY, VXB, VYB, P, F, VX_ACCUM, VY_ACCUM, R, VX_CURRENT, VY_CURRENT = range(12)
ST_ACCUM, ST_DIFFUSE, ST_ADVECT, ST_SETVBCS, ST_PROJECT, ST_SAVE = range(6)
BC_INFLOW, BC_OUTFLOW, BC_OUTFLOW_ZERO_PRESSURE, BC_NOSLIP = range(4)


def build(ref=True):
    """make the C port and, when /root/reference is here, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "_build/liboracle.so"])
    if ref and os.path.exists("/root/reference/simulation.cpp"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def fp(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(FP)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def field_shape(field, w, h):
    if field in (VX, VXB, VX_ACCUM, VX_CURRENT):
        return (h, w - 1)
    if field in (VY, VYB, VY_ACCUM, VY_CURRENT):
        return (h - 1, w)
    return (h, w)


def mg_level_sizes(w, h):
    """pressure_solver.hpp:20-29: halve by integer division while w>3 && h>3."""
    out = []
    while w > 3 and h > 3:
        out.append((w, h))
        w //= 2
        h //= 2
    return out


UP = C.POINTER(C.c_uint)
# CoItem + CoKinematicsSimple (components.hpp:6-43) == orc_item == ubgl_item
ITEM_DTYPE = np.dtype([("size", np.float32, 2), ("pos", np.float32, 2), ("rotation", np.float32),
                       ("mass", np.float32), ("vel", np.float32, 2), ("force", np.float32, 2),
                       ("angVel", np.float32), ("angForce", np.float32), ("bumpCount", np.int32)])
assert ITEM_DTYPE.itemsize == 52


class _SimBase:
    """Shared python face of a CPU Simulation handle (ref or port)."""

    prefix = None

    def __init__(self, lib, flag, pwidth=0.8, mu=0.001):
        self.lib = lib
        flag = f32(flag)
        self.H, self.W = flag.shape
        self.h = getattr(lib, self.prefix + "sim_create")(fp(flag), self.W, self.H, pwidth, mu)
        self.dx = pwidth / (self.W - 1.0)

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def close(self):
        if self.h:
            self._fn("sim_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get(self, field):
        a = np.empty(field_shape(field, self.W, self.H), np.float32)
        assert self._fn("sim_get")(self.h, field, fp(a)) == 0
        return a

    def set(self, field, a):
        a = f32(a)
        assert a.shape == field_shape(field, self.W, self.H), (a.shape, field)
        assert self._fn("sim_set")(self.h, field, fp(a)) == 0

    def update_flag(self, flag):
        self._fn("sim_update_flag")(self.h, fp(f32(flag)))

    def set_bc(self, west, east, north, south):
        self._fn("sim_set_bc")(self.h, west, east, north, south)

    def add_sink(self, x, y, z):
        self._fn("sim_add_sink")(self.h, x, y, z)

    def sinks(self):
        n = self._fn("sim_num_sinks")(self.h)
        a = np.zeros((n, 3), np.float32)
        if n:
            self._fn("sim_get_sinks")(self.h, fp(a))
        return a

    def step(self, dt):
        self._fn("sim_step")(self.h, dt)

    def stage(self, stage, dt):
        self._fn("sim_stage")(self.h, stage, dt)

    def mg_levels(self):
        return self._fn("sim_mg_levels")(self.h)

    def mg_flagc(self, level):
        w, h = C.c_int(), C.c_int()
        self._fn("sim_mg_level_size")(self.h, level, C.byref(w), C.byref(h))
        a = np.empty((h.value, w.value), np.float32)
        self._fn("sim_mg_get_flagc")(self.h, level, fp(a))
        return a


class _MGBase:
    prefix = None

    def __init__(self, lib, w, h):
        self.lib = lib
        self.W, self.H = w, h
        self.h = getattr(lib, self.prefix + "mg_create")(w, h)

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def close(self):
        if self.h:
            self._fn("mg_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def levels(self):
        return self._fn("mg_levels")(self.h)

    def update_fields(self, flag):
        self._fn("mg_update_fields")(self.h, fp(f32(flag)))

    def flagc(self, level):
        w, h = C.c_int(), C.c_int()
        self._fn("mg_level_size")(self.h, level, C.byref(w), C.byref(h))
        a = np.empty((h.value, w.value), np.float32)
        self._fn("mg_get_flagc")(self.h, level, fp(a))
        return a

    def set(self, p=None, f=None, flag=None):
        args = [fp(f32(a)) if a is not None else None for a in (p, f, flag)]
        self._keep = args
        self._fn("mg_set")(self.h, *args)

    def get_p(self):
        a = np.empty((self.H, self.W), np.float32)
        self._fn("mg_get_p")(self.h, fp(a))
        return a

    def solve(self, hh, zero_gradient_bc=False):
        self._fn("mg_solve")(self.h, hh, int(zero_gradient_bc))

    def residual(self, hh):
        return float(self._fn("mg_residual")(self.h, hh))


def _declare(lib, pre):
    v, i, f = C.c_void_p, C.c_int, C.c_float
    sig = {
        "rbgs": (None, [FP, FP, FP, i, i, f, f, i]),
        "residual": (f, [FP, FP, FP, FP, i, i, f]),
        "restrict": (None, [FP, i, i, FP, i, i]),
        "prolongate": (None, [FP, i, i, FP, FP, i, i, FP]),
        "correct": (None, [FP, FP, i, i]),
        "zero_gradient_bc": (None, [FP, i, i]),
        "mg_create": (v, [i, i]),
        "mg_destroy": (None, [v]),
        "mg_levels": (i, [v]),
        "mg_level_size": (None, [v, i, IP, IP]),
        "mg_update_fields": (None, [v, FP]),
        "mg_get_flagc": (None, [v, i, FP]),
        "mg_set": (None, [v, FP, FP, FP]),
        "mg_get_p": (None, [v, FP]),
        "mg_solve": (None, [v, f, i]),
        "mg_residual": (f, [v, f]),
        "sim_create": (v, [FP, i, i, f, f]),
        "sim_destroy": (None, [v]),
        "sim_get": (i, [v, i, FP]),
        "sim_set": (i, [v, i, FP]),
        "sim_update_flag": (None, [v, FP]),
        "sim_set_bc": (None, [v, i, i, i, i]),
        "sim_add_sink": (None, [v, f, f, f]),
        "sim_num_sinks": (i, [v]),
        "sim_get_sinks": (None, [v, FP]),
        "sim_h": (f, [v]),
        "sim_step": (None, [v, f]),
        "sim_stage": (None, [v, i, f]),
        "sim_mg_levels": (i, [v]),
        "sim_mg_level_size": (None, [v, i, IP, IP]),
        "sim_mg_get_flagc": (None, [v, i, FP]),
        "set_threads": (None, [i]),
        "max_threads": (i, []),
        "num_procs": (i, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, pre + name)
        fn.restype = res
        fn.argtypes = args


class _Checker:
    """Stateless stage functions + factories, same face for Ref and Port."""

    prefix = None
    so = None

    def __init__(self):
        if not os.path.exists(self.so):
            raise FileNotFoundError(self.so)
        self.lib = C.CDLL(self.so)
        _declare(self.lib, self.prefix)
        sim_cls = type("Sim", (_SimBase,), {"prefix": self.prefix})
        mg_cls = type("MG", (_MGBase,), {"prefix": self.prefix})
        self.Sim = lambda flag, pwidth=0.8, mu=0.001: sim_cls(self.lib, flag, pwidth, mu)
        self.MG = lambda w, h: mg_cls(self.lib, w, h)

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def set_threads(self, n):
        self._fn("set_threads")(int(n))

    def max_threads(self):
        return self._fn("max_threads")()

    def num_procs(self):
        return self._fn("num_procs")()

    def canonical_threads(self, height):
        """rbgs takes the canonical red-black path iff height/T < 100
        (pressure_solver.cpp:65): pin T so every level qualifies."""
        self.set_threads(max(2, height // 100 + 1))

    def rbgs(self, p, f, flag, hh, alpha=1.0, sweeps=1):
        p = f32(p).copy()
        H, W = p.shape
        self._fn("rbgs")(fp(p), fp(f32(f)), fp(f32(flag)), W, H, hh, alpha, sweeps)
        return p

    def residual(self, p, f, flag, hh):
        p = f32(p)
        H, W = p.shape
        r = np.zeros_like(p)
        l2 = self._fn("residual")(fp(p), fp(f32(f)), fp(f32(flag)), fp(r), W, H, hh)
        return r, float(l2)

    def restrict(self, r):
        r = f32(r)
        H, W = r.shape
        rc = np.zeros((H // 2, W // 2), np.float32)
        self._fn("restrict")(fp(r), W, H, fp(rc), W // 2, H // 2)
        return rc

    def prolongate(self, ec, flagc, flag):
        flag = f32(flag)
        H, W = flag.shape
        ec = f32(ec)
        hc, wc = ec.shape
        e = np.full((H, W), 7.0, np.float32)  # prolongate() zero-fills first
        self._fn("prolongate")(fp(e), W, H, fp(ec), fp(f32(flagc)), wc, hc, fp(flag))
        return e

    def correct(self, p, e):
        p = f32(p).copy()
        H, W = p.shape
        self._fn("correct")(fp(p), fp(f32(e)), W, H)
        return p

    def zero_gradient_bc(self, p):
        p = f32(p).copy()
        H, W = p.shape
        self._fn("zero_gradient_bc")(fp(p), W, H)
        return p


class Ref(_Checker):
    prefix = "ref_"
    so = REF_SO

    def __init__(self):
        super().__init__()
        v, i, f = C.c_void_p, C.c_int, C.c_float
        L = self.lib
        L.ref_terrain_create.restype = v
        L.ref_terrain_create.argtypes = [C.c_char_p, i]
        L.ref_terrain_destroy.argtypes = [v]
        L.ref_terrain_size.argtypes = [v, IP, IP]
        L.ref_terrain_flag.argtypes = [v, FP]
        L.ref_terrain_draw_circle.argtypes = [v, f, f, i, f]
        L.ref_items_advect_simple.argtypes = [v, v, i, f]
        L.ref_items_view_order.argtypes = [i, IP]
        if hasattr(L, "ref_items_advect"):
            L.ref_items_advect.argtypes = [v, v, i, f]

    def items_advect_simple(self, sim, items, game_dt):
        """Unmodified Simulation::advectFloatingItemsSimple on sim's fields (in place on items)."""
        assert items.dtype == ITEM_DTYPE and items.flags["C_CONTIGUOUS"]
        self.lib.ref_items_advect_simple(sim.h, items.ctypes.data_as(C.c_void_p), len(items), game_dt)

    def items_advect(self, sim, items, game_dt):
        """Unmodified Simulation::advectFloatingItems (rigid bodies) on sim's fields (in place on items)."""
        assert items.dtype == ITEM_DTYPE and items.flags["C_CONTIGUOUS"]
        self.lib.ref_items_advect(sim.h, items.ctypes.data_as(C.c_void_p), len(items), game_dt)

    def items_view_order(self, n):
        o = np.zeros(n, np.int32)
        self.lib.ref_items_view_order(n, o.ctypes.data_as(IP))
        return o

    def terrain(self, png, scale=1):
        return _RefTerrain(self.lib, png, scale)

    def terrain_flag(self, png, scale=1):
        t = self.lib.ref_terrain_create(png.encode(), scale)
        w, h = C.c_int(), C.c_int()
        self.lib.ref_terrain_size(t, C.byref(w), C.byref(h))
        a = np.empty((h.value, w.value), np.float32)
        self.lib.ref_terrain_flag(t, fp(a))
        self.lib.ref_terrain_destroy(t)
        return a


class _RefTerrain:
    """The reference's Terrain object (terrain.hpp:9-32) for drawCircle fixtures."""

    def __init__(self, lib, png, scale):
        self.lib = lib
        self.t = lib.ref_terrain_create(png.encode(), scale)

    def flag(self):
        w, h = C.c_int(), C.c_int()
        self.lib.ref_terrain_size(self.t, C.byref(w), C.byref(h))
        a = np.empty((h.value, w.value), np.float32)
        self.lib.ref_terrain_flag(self.t, fp(a))
        return a

    def draw_circle(self, x, y, diam, val):
        self.lib.ref_terrain_draw_circle(self.t, x, y, diam, val)

    def __del__(self):
        try:
            self.lib.ref_terrain_destroy(self.t)
        except Exception:
            pass


class Port(_Checker):
    prefix = "orc_"
    so = PORT_SO

    def __init__(self):
        super().__init__()
        v, i, f, u = C.c_void_p, C.c_int, C.c_float, C.c_uint
        L = self.lib
        L.orc_colocate.argtypes = [FP, FP, i, i, FP, FP]
        L.orc_tracers_advect.argtypes = [FP, UP, UP, FP, i, i, f, f, f, u, FP, i, i, FP, i, i]
        L.orc_tracers_shift.argtypes = [FP, i, i, f]
        L.orc_items_advect_simple.argtypes = [v, i, f, FP, FP, FP, FP, FP, FP, i, i, f]
        L.orc_items_advect.argtypes = [v, i, f, FP, FP, FP, FP, FP, i, i, f]
        L.orc_draw_circle.argtypes = [FP, FP, i, i, f, f, i, f]
        L.orc_set_grids_all.argtypes = [FP, FP, FP, FP, FP, i, i]
        L.orc_shift_map.argtypes = [FP] * 9 + [i, i]

    # ---- SURVEY.md 8f "next" rows (oracle/ubgl_oracle_next.c) ----
    def colocate(self, vx, vy):
        """interp_shader.cs: staggered vx (H x W-1), vy (H-1 x W) -> vxy (2H-1, 2W-1, 2), mag."""
        ny, nx = vx.shape[0], vy.shape[1]
        vxy = np.empty((2 * ny - 1, 2 * nx - 1, 2), np.float32)
        mag = np.empty((2 * ny - 1, 2 * nx - 1), np.float32)
        self.lib.orc_colocate(fp(f32(vx)), fp(f32(vy)), nx, ny, fp(vxy), fp(mag))
        return vxy, mag

    def tracers_advect(self, st, dt, pdim, rand_seed, vxy, flagtex):
        """advect_tracer_points.cs on st = dict(points (nt, np, 2) f32, start, end u32, ages f32), in place."""
        nt, npts = st["points"].shape[:2]
        th, tw = vxy.shape[:2]
        fh, fw = flagtex.shape
        self.lib.orc_tracers_advect(fp(st["points"]), st["start"].ctypes.data_as(UP),
                                    st["end"].ctypes.data_as(UP), fp(st["ages"]), nt, npts, dt,
                                    pdim[0], pdim[1], rand_seed & 0xFFFFFFFF, fp(f32(vxy)), tw, th,
                                    fp(f32(flagtex)), fw, fh)

    def tracers_shift(self, st, shift):
        nt, npts = st["points"].shape[:2]
        self.lib.orc_tracers_shift(fp(st["points"]), nt, npts, shift)

    def items_advect_simple(self, items, game_dt, flag, vx, vy, p, vx_accum, vy_accum, pwidth=0.8):
        """advectFloatingItemsSimple; items and the accumulators are updated in place."""
        assert items.dtype == ITEM_DTYPE and items.flags["C_CONTIGUOUS"]
        H, W = flag.shape
        self.lib.orc_items_advect_simple(items.ctypes.data_as(C.c_void_p), len(items), game_dt,
                                         fp(f32(flag)), fp(f32(vx)), fp(f32(vy)), fp(f32(p)),
                                         fp(vx_accum), fp(vy_accum), W, H, pwidth)

    def items_advect(self, items, game_dt, flag, vx, vy, vx_accum, vy_accum, pwidth=0.8):
        """advectFloatingItems (rigid bodies); items and the accumulators are updated in place."""
        assert items.dtype == ITEM_DTYPE and items.flags["C_CONTIGUOUS"]
        H, W = flag.shape
        self.lib.orc_items_advect(items.ctypes.data_as(C.c_void_p), len(items), game_dt, fp(f32(flag)),
                                  fp(f32(vx)), fp(f32(vy)), fp(vx_accum), fp(vy_accum), W, H, pwidth)

    def draw_circle(self, flag_full, flag_sim, cx, cy, diam, val):
        h, w = flag_sim.shape
        self.lib.orc_draw_circle(fp(flag_full), fp(flag_sim), w, h, cx, cy, diam, val)

    def set_grids_all(self, flag, vx, vy, p, newflag):
        H, W = flag.shape
        self.lib.orc_set_grids_all(fp(flag), fp(vx), fp(vy), fp(p), fp(f32(newflag)), W, H)

    def shift_map(self, flag, vxf, vxb, vyf, vyb, p, vxc, vyc, newflag):
        H, W = flag.shape
        self.lib.orc_shift_map(fp(flag), fp(vxf), fp(vxb), fp(vyf), fp(vyb), fp(p), fp(vxc), fp(vyc),
                               fp(f32(newflag)), W, H)


class RefStrict(_Checker):
    """The unmodified reference fluid TUs compiled -O2 -ffp-contract=off instead of -Ofast
    (oracle/Makefile): what the reference computes under IEEE-strict code generation.  Its
    distance from ``Ref`` is the reference's own rounding sensitivity."""
    prefix = "ref_"
    so = REF_STRICT_SO


def have_ref():
    return os.path.exists(REF_SO)


def have_ref_strict():
    return os.path.exists(REF_STRICT_SO)


def have_port():
    return os.path.exists(PORT_SO)


GLSL_SO = os.path.join(HERE, "_ref", "libubgl_glsl.so")


def have_glsl():
    return os.path.exists(GLSL_SO)


class Glsl:
    """oracle/_ref/libubgl_glsl.so: the reference's UNMODIFIED GLSL compute shaders (interp_shader.cs,
    advect_tracer_points.cs) compiled as C++ through oracle/shim/glsl_shim.hpp and dispatched as the
    reference's host code does (oracle/glsl_run.cpp).  Same call signatures as Port.colocate /
    Port.tracers_advect, which it pins."""

    def __init__(self):
        self.lib = C.CDLL(GLSL_SO)
        self.lib.glsl_colocate.argtypes = [FP, FP, C.c_int, C.c_int, FP, FP]
        self.lib.glsl_colocate.restype = None
        self.lib.glsl_tracers_advect.argtypes = [FP, UP, UP, FP, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                                 C.c_uint, FP, C.c_int, C.c_int, FP, C.c_int, C.c_int]
        self.lib.glsl_tracers_advect.restype = None
        self.lib.glsl_tracers_shift.argtypes = [FP, FP, C.c_int, C.c_int, C.c_float]
        self.lib.glsl_tracers_shift.restype = None

    def tracers_shift(self, st, shift):
        """shift_tracers.cs; the ribbon vertices it also moves are rendering state (a scratch array here)."""
        nt, npts = st["points"].shape[:2]
        vertices = np.zeros((nt * npts * 2, 2), np.float32)
        self.lib.glsl_tracers_shift(fp(st["points"]), fp(vertices), nt, npts, shift)

    def colocate(self, vx, vy):
        ny, nx = vx.shape[0], vy.shape[1]
        vxy = np.empty((2 * ny - 1, 2 * nx - 1, 2), np.float32)
        mag = np.empty((2 * ny - 1, 2 * nx - 1), np.float32)
        self.lib.glsl_colocate(fp(f32(vx)), fp(f32(vy)), nx, ny, fp(vxy), fp(mag))
        return vxy, mag

    def tracers_advect(self, st, dt, pdim, rand_seed, vxy, flagtex):
        nt, npts = st["points"].shape[:2]
        th, tw = vxy.shape[:2]
        fh, fw = flagtex.shape
        self.lib.glsl_tracers_advect(fp(st["points"]), st["start"].ctypes.data_as(UP), st["end"].ctypes.data_as(UP),
                                     fp(st["ages"]), nt, npts, dt, pdim[0], pdim[1], rand_seed & 0xFFFFFFFF,
                                     fp(f32(vxy)), tw, th, fp(f32(flagtex)), fw, fh)
