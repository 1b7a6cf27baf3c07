"""ctypes binding of include/ubgl.h, shaped like the reference's classes.

``Simulation`` mirrors simulation.hpp:18-141 (step, setVBCs..., public grids as
get/set of numpy arrays), ``MG`` mirrors pressure_solver.hpp:13-76.  Every call
goes through libubgl.so; there is no Python or CPU implementation behind it.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UBGL_LIB_PATH") or os.path.join(HERE, "_lib", "libubgl.so")  # override: A/B builds only

FLAG, VX, VY, VXB, VYB, P, F, VX_ACCUM, VY_ACCUM, R, VX_CURRENT, VY_CURRENT = range(12)
ST_ACCUM, ST_DIFFUSE, ST_ADVECT, ST_SETVBCS, ST_PROJECT, ST_SAVE = range(6)
BC_INFLOW, BC_OUTFLOW, BC_OUTFLOW_ZERO_PRESSURE, BC_NOSLIP = range(4)
OPT_VCYCLES, OPT_FUSED, OPT_GRAPH, OPT_TIMING = range(4)

FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int)
UP = C.POINTER(C.c_uint)
# struct ubgl_item (include/ubgl.h) = CoItem + CoKinematicsSimple (components.hpp:6-43)
ITEM_DTYPE = np.dtype([("size", np.float32, 2), ("pos", np.float32, 2), ("rotation", np.float32),
                       ("mass", np.float32), ("vel", np.float32, 2), ("force", np.float32, 2),
                       ("angVel", np.float32), ("angForce", np.float32), ("bumpCount", np.int32)])


class UbglError(RuntimeError):
    pass


class HostMirrors(C.Structure):
    _fields_ = [("flag", FP), ("vx_accum", FP), ("vy_accum", FP), ("vx", FP), ("vy", FP),
                ("p", FP), ("vx_current", FP), ("vy_current", FP)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise UbglError(
            f"{LIB_PATH} is missing: build it with `make -C ubootgl_b200` "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    v, i, f, ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
    VP = C.POINTER(C.c_void_p)
    sig = {
        "ubgl_version": (i, []),
        "ubgl_last_error": (C.c_char_p, []),
        "ubgl_device_count": (i, []),
        "ubgl_sim_create": (i, [FP, i, i, f, f, i, VP]),
        "ubgl_sim_destroy": (i, [v]),
        "ubgl_sim_set_option": (i, [v, i, i]),
        "ubgl_sim_set_bc": (i, [v, i, i, i, i]),
        "ubgl_sim_upload": (i, [v, i, FP]),
        "ubgl_sim_upload_add": (i, [v, i, FP]),
        "ubgl_sim_download": (i, [v, i, FP]),
        "ubgl_sim_update_flag": (i, [v, FP]),
        "ubgl_sim_mg_levels": (i, [v]),
        "ubgl_sim_mg_level_size": (i, [v, i, IP, IP]),
        "ubgl_sim_mg_get_flagc": (i, [v, i, FP]),
        "ubgl_sim_set_sinks": (i, [v, FP, i]),
        "ubgl_sim_get_sinks": (i, [v, FP, i, IP]),
        "ubgl_sim_step": (i, [v, f]),
        "ubgl_sim_stage": (i, [v, i, f]),
        "ubgl_sim_step_host": (i, [v, f, C.POINTER(HostMirrors)]),
        "ubgl_sim_step_host_pipelined": (i, [v, f, C.POINTER(HostMirrors)]),
        "ubgl_sim_step_host_flush": (i, [v, C.POINTER(HostMirrors)]),
        "ubgl_sim_set_tolerance": (i, [v, f, i, f]),
        "ubgl_sim_solve_info": (i, [v, IP, FP, FP, i, IP]),
        "ubgl_sim_sync": (i, [v]),
        "ubgl_sim_residual": (i, [v, FP]),
        "ubgl_sim_mg_solve": (i, [v, i]),
        "ubgl_sim_mg_solve_ex": (i, [v, f, i, i]),
        "ubgl_sim_device_ptr": (i, [v, i, VP, IP]),
        "ubgl_sim_stage_ms": (i, [v, i, FP]),
        "ubgl_sim_launch_count": (ll, [v]),
        "ubgl_sim_stream": (v, [v]),
        "ubgl_sim_profile": (i, [v, i]),
        "ubgl_num_kernel_kinds": (i, []),
        "ubgl_kernel_kind_name": (C.c_char_p, [i]),
        "ubgl_sim_kernel_stats": (i, [v, i, i, C.POINTER(ll), C.POINTER(C.c_double)]),
        "ubgl_mg_create": (i, [i, i, i, VP]),
        "ubgl_mg_destroy": (i, [v]),
        "ubgl_mg_set_option": (i, [v, i, i]),
        "ubgl_mg_levels": (i, [v]),
        "ubgl_mg_level_size": (i, [v, i, IP, IP]),
        "ubgl_mg_update_fields": (i, [v, FP]),
        "ubgl_mg_get_flagc": (i, [v, i, FP]),
        "ubgl_mg_solve_host": (i, [v, FP, FP, FP, f, i]),
        "ubgl_mg_upload": (i, [v, FP, FP, FP]),
        "ubgl_mg_download_p": (i, [v, FP]),
        "ubgl_mg_solve": (i, [v, f, i, i]),
        "ubgl_mg_residual": (i, [v, f, FP]),
        "ubgl_mg_sync": (i, [v]),
        "ubgl_mg_launch_count": (ll, [v]),
        "ubgl_mg_stream": (v, [v]),
        "ubgl_slab_plan": (i, [i, i, i, i, IP, IP, IP, IP, IP, IP]),
        "ubgl_slab_create": (i, [FP, i, i, f, f, i, i, i, VP]),
        "ubgl_slab_destroy": (i, [v]),
        "ubgl_slab_ipc_size": (i, []),
        "ubgl_slab_ipc_export": (i, [v, v]),
        "ubgl_slab_connect": (i, [v, v]),
        "ubgl_slab_field_rows": (i, [v, i, IP, IP, IP]),
        "ubgl_slab_upload": (i, [v, i, FP]),
        "ubgl_slab_download": (i, [v, i, FP]),
        "ubgl_slab_set_option": (i, [v, i, i]),
        "ubgl_slab_set_sinks": (i, [v, FP, i]),
        "ubgl_slab_step": (i, [v, f]),
        "ubgl_slab_set_row_weights": (i, [FP, i]),
        "ubgl_slab_step_host": (i, [v, f, C.POINTER(HostMirrors)]),
        "ubgl_slab_sync": (i, [v]),
        "ubgl_slab_residual_sumsq": (i, [v, C.POINTER(C.c_double)]),
        "ubgl_slab_launch_count": (ll, [v]),
        "ubgl_slab_stream": (v, [v]),
        "ubgl_slab_stats": (i, [v, C.POINTER(ll), C.POINTER(ll)]),
        "ubgl_slab_profile": (i, [v, i]),
        "ubgl_slab_kernel_stats": (i, [v, i, i, C.POINTER(ll), C.POINTER(C.c_double)]),
        "ubgl_tracers_create": (i, [i, i, i, VP]),
        "ubgl_tracers_destroy": (i, [v]),
        "ubgl_tracers_upload": (i, [v, FP, UP, UP, FP]),
        "ubgl_tracers_download": (i, [v, FP, UP, UP, FP]),
        "ubgl_tracers_set_flag_texture": (i, [v, FP, i, i]),
        "ubgl_tracers_advect": (i, [v, v, f, C.c_uint]),
        "ubgl_tracers_shift": (i, [v, f]),
        "ubgl_sim_colocate_velocity": (i, [v, FP, FP]),
        "ubgl_sim_export_display": (i, [v, v, v, v]),
        "ubgl_display_array_create": (i, [i, i, i, i, VP]),
        "ubgl_display_array_read": (i, [v, FP]),
        "ubgl_display_array_destroy": (i, [v]),
        "ubgl_items_create": (i, [i, VP]),
        "ubgl_items_destroy": (i, [v]),
        "ubgl_items_upload": (i, [v, v, i]),
        "ubgl_items_download": (i, [v, v, i, IP]),
        "ubgl_items_advect_simple": (i, [v, v, f]),
        "ubgl_items_advect": (i, [v, v, f]),
        "ubgl_sim_draw_circles": (i, [v, FP, i, f]),
        "ubgl_sim_set_grids_all": (i, [v, FP]),
        "ubgl_sim_shift_map": (i, [v, FP]),
        "ubgl_rbgs": (i, [FP, FP, FP, i, i, f, f, i]),
        "ubgl_residual": (i, [FP, FP, FP, FP, i, i, f, FP]),
        "ubgl_restrict": (i, [FP, i, i, FP]),
        "ubgl_prolongate": (i, [FP, i, i, FP, FP, FP]),
        "ubgl_correct": (i, [FP, FP, i, i]),
        "ubgl_zero_gradient_bc": (i, [FP, i, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._ubgl_sig = sig
    return L


lib = _load()
SYMBOLS = sorted(lib._ubgl_sig)


def _ck(rc):
    if rc != 0:
        raise UbglError(f"libubgl error {rc}: {lib.ubgl_last_error().decode()}")


def _fp(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(FP)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def field_shape(field, w, h):
    if field in (VX, VXB, VX_ACCUM, VX_CURRENT):
        return (h, w - 1)
    if field in (VY, VYB, VY_ACCUM, VY_CURRENT):
        return (h - 1, w)
    return (h, w)


class Simulation:
    """Simulation(flag, pwidth, mu) -- simulation.hpp:32-67, state on the GPU."""

    def __init__(self, flag, pwidth=0.8, mu=0.001, device=0):
        flag = _f32(flag)
        self.height, self.width = flag.shape
        self.pwidth, self.mu = pwidth, mu
        self.h = np.float32(pwidth) / (np.float32(self.width) - np.float32(1.0))
        self._h = C.c_void_p()
        _ck(lib.ubgl_sim_create(_fp(flag), self.width, self.height, pwidth, mu, device,
                                C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib.ubgl_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- public members as arrays ----
    def get(self, field):
        a = np.empty(field_shape(field, self.width, self.height), np.float32)
        _ck(lib.ubgl_sim_download(self._h, field, _fp(a)))
        return a

    def set(self, field, a):
        a = _f32(a)
        if a.shape != field_shape(field, self.width, self.height):
            raise UbglError(f"field {field}: shape {a.shape} does not match the grid")
        _ck(lib.ubgl_sim_upload(self._h, field, _fp(a)))

    def add(self, field, a):
        """field += a (host contributions to a device-resident accumulator)."""
        a = _f32(a)
        if a.shape != field_shape(field, self.width, self.height):
            raise UbglError(f"field {field}: shape {a.shape} does not match the grid")
        _ck(lib.ubgl_sim_upload_add(self._h, field, _fp(a)))

    def update_flag(self, flag):
        """memcpy into sim.flag + mg.updateFields (ubootgl_app.cpp:111-112)."""
        flag = _f32(flag)
        if flag.shape != (self.height, self.width):
            raise UbglError("flag shape does not match the grid")
        _ck(lib.ubgl_sim_update_flag(self._h, _fp(flag)))

    def set_bc(self, west, east, north, south):
        _ck(lib.ubgl_sim_set_bc(self._h, west, east, north, south))

    def set_option(self, opt, val):
        _ck(lib.ubgl_sim_set_option(self._h, opt, int(val)))

    def set_sinks(self, xyz):
        xyz = _f32(np.asarray(xyz, np.float32).reshape(-1, 3))
        _ck(lib.ubgl_sim_set_sinks(self._h, _fp(xyz), len(xyz)))

    def add_sink(self, x, y, z):
        s = self.sinks()
        self.set_sinks(np.concatenate([s, np.array([[x, y, z]], np.float32)]))

    def sinks(self):
        n = C.c_int()
        _ck(lib.ubgl_sim_get_sinks(self._h, None, 0, C.byref(n)))
        a = np.zeros((n.value, 3), np.float32)
        if n.value:
            _ck(lib.ubgl_sim_get_sinks(self._h, _fp(a), n.value, C.byref(n)))
        return a

    # ---- methods ----
    def step(self, dt):
        _ck(lib.ubgl_sim_step(self._h, dt))

    def stage(self, stage, dt):
        _ck(lib.ubgl_sim_stage(self._h, stage, dt))

    def step_host(self, dt, flag=None, vx_accum=None, vy_accum=None, vx=None, vy=None, p=None,
                  vx_current=None, vy_current=None):
        m = HostMirrors()
        for name, a in (("flag", flag), ("vx_accum", vx_accum), ("vy_accum", vy_accum),
                        ("vx", vx), ("vy", vy), ("p", p), ("vx_current", vx_current),
                        ("vy_current", vy_current)):
            setattr(m, name, _fp(a) if a is not None else None)
        _ck(lib.ubgl_sim_step_host(self._h, dt, C.byref(m)))

    @staticmethod
    def _mirrors(**kw):
        m = HostMirrors()
        for name in ("flag", "vx_accum", "vy_accum", "vx", "vy", "p", "vx_current", "vy_current"):
            a = kw.get(name)
            setattr(m, name, _fp(a) if a is not None else None)
        return m

    def step_host_pipelined(self, dt, **mirrors):
        """ubgl_sim_step_host_pipelined: the mirrors written are those of the PREVIOUS step."""
        m = self._mirrors(**mirrors)
        _ck(lib.ubgl_sim_step_host_pipelined(self._h, dt, C.byref(m)))

    def step_host_flush(self, **mirrors):
        m = self._mirrors(**mirrors)
        _ck(lib.ubgl_sim_step_host_flush(self._h, C.byref(m)))

    def sync(self):
        _ck(lib.ubgl_sim_sync(self._h))

    def set_tolerance(self, rel_tol, max_cycles=20, stagnation=0.9):
        """rel_tol <= 0: the reference's fixed 2 V-cycles (simulation.cpp:189-190)."""
        _ck(lib.ubgl_sim_set_tolerance(self._h, rel_tol, max_cycles, stagnation))

    def solve_info(self):
        """(cycles_done, ||f*flag||, [||r|| before the first / after every cycle])."""
        n, c, fn = C.c_int(), C.c_int(), C.c_float()
        hist = np.zeros(256, np.float32)
        _ck(lib.ubgl_sim_solve_info(self._h, C.byref(c), C.byref(fn), _fp(hist), 256, C.byref(n)))
        return c.value, fn.value, hist[:min(n.value, 256)].copy()

    def residual(self):
        l2 = C.c_float()
        _ck(lib.ubgl_sim_residual(self._h, C.byref(l2)))
        return l2.value

    def mg_solve(self, cycles=1):
        _ck(lib.ubgl_sim_mg_solve(self._h, cycles))

    def mg_levels(self):
        return lib.ubgl_sim_mg_levels(self._h)

    def mg_flagc(self, level):
        w, h = C.c_int(), C.c_int()
        _ck(lib.ubgl_sim_mg_level_size(self._h, level, C.byref(w), C.byref(h)))
        a = np.empty((h.value, w.value), np.float32)
        _ck(lib.ubgl_sim_mg_get_flagc(self._h, level, _fp(a)))
        return a

    def device_ptr(self, field):
        p, pitch = C.c_void_p(), C.c_int()
        _ck(lib.ubgl_sim_device_ptr(self._h, field, C.byref(p), C.byref(pitch)))
        return p.value, pitch.value

    def stage_ms(self, stage):
        ms = C.c_float()
        _ck(lib.ubgl_sim_stage_ms(self._h, stage, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return lib.ubgl_sim_launch_count(self._h)

    def profile(self, on):
        _ck(lib.ubgl_sim_profile(self._h, int(on)))

    def kernel_stats(self):
        """{(kind_name, level): (launches, total_ms)} of the profiled region."""
        out = {}
        for k in range(lib.ubgl_num_kernel_kinds()):
            for l in range(16):
                n, ms = C.c_longlong(), C.c_double()
                _ck(lib.ubgl_sim_kernel_stats(self._h, k, l, C.byref(n), C.byref(ms)))
                if n.value:
                    out[(lib.ubgl_kernel_kind_name(k).decode(), l)] = (n.value, ms.value)
        return out

    def stream(self):
        return lib.ubgl_sim_stream(self._h)

    # ---- callers either side of the step, on the device ----
    def colocate_velocity(self, download=True):
        """interp_shader.cs: (vxy (2H-1, 2W-1, 2), mag (2H-1, 2W-1)) from vx_current / vy_current."""
        if not download:
            _ck(lib.ubgl_sim_colocate_velocity(self._h, None, None))
            return None
        th, tw = 2 * self.height - 1, 2 * self.width - 1
        vxy, mag = np.empty((th, tw, 2), np.float32), np.empty((th, tw), np.float32)
        _ck(lib.ubgl_sim_colocate_velocity(self._h, _fp(vxy), _fp(mag)))
        return vxy, mag

    def export_display(self, vxy=None, mag=None, p=None):
        """velocity_textures.cpp:63-93 + draw_2dbuf.cpp:181-209 without PCIe: texels into CUDA arrays
        (mapped GL textures, or DisplayArray stand-ins)."""
        h = lambda a: a._h if isinstance(a, DisplayArray) else a
        _ck(lib.ubgl_sim_export_display(self._h, h(vxy), h(mag), h(p)))

    def draw_circles(self, xyd, val):
        """Terrain::drawCircle (terrain.cpp:213-234) x n + MG::updateFields, on the device."""
        xyd = _f32(np.asarray(xyd, np.float32).reshape(-1, 3))
        _ck(lib.ubgl_sim_draw_circles(self._h, _fp(xyd), len(xyd), val))

    def set_grids_all(self, newflag=None):
        _ck(lib.ubgl_sim_set_grids_all(self._h, _fp(_f32(newflag)) if newflag is not None else None))

    def shift_map(self, new_last_column):
        col = _f32(new_last_column)
        if col.shape != (self.height,):
            raise UbglError("shift_map: the new column needs H values")
        _ck(lib.ubgl_sim_shift_map(self._h, _fp(col)))


class DisplayArray:
    """A w x h CUDA array of 1 or 2 float channels: what a mapped R32F / RG32F GL texture is to CUDA."""

    def __init__(self, w, h, channels, device=0):
        self.w, self.h, self.channels = w, h, channels
        self._h = C.c_void_p()
        _ck(lib.ubgl_display_array_create(w, h, channels, device, C.byref(self._h)))

    def read(self):
        out = np.empty((self.h, self.w, self.channels) if self.channels > 1 else (self.h, self.w), np.float32)
        _ck(lib.ubgl_display_array_read(self._h, _fp(out)))
        return out

    def close(self):
        if self._h:
            lib.ubgl_display_array_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Tracers:
    """DrawTracersCS::GLTracers (draw_tracers_cs.cpp:27-98) + advect_tracer_points.cs on the GPU."""

    def __init__(self, ntracers=1000, npoints=30, device=0):
        self.ntracers, self.npoints = ntracers, npoints
        self._h = C.c_void_p()
        _ck(lib.ubgl_tracers_create(ntracers, npoints, device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib.ubgl_tracers_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, st):
        _ck(lib.ubgl_tracers_upload(self._h, _fp(st["points"]), st["start"].ctypes.data_as(UP),
                                    st["end"].ctypes.data_as(UP), _fp(st["ages"])))

    def state(self):
        st = dict(points=np.empty((self.ntracers, self.npoints, 2), np.float32),
                  start=np.empty(self.ntracers, np.uint32), end=np.empty(self.ntracers, np.uint32),
                  ages=np.empty(self.ntracers, np.float32))
        _ck(lib.ubgl_tracers_download(self._h, _fp(st["points"]), st["start"].ctypes.data_as(UP),
                                      st["end"].ctypes.data_as(UP), _fp(st["ages"])))
        return st

    def set_flag_texture(self, flag):
        if flag is None:
            _ck(lib.ubgl_tracers_set_flag_texture(self._h, None, 0, 0))
        else:
            flag = _f32(flag)
            _ck(lib.ubgl_tracers_set_flag_texture(self._h, _fp(flag), flag.shape[1], flag.shape[0]))

    def advect(self, sim, dt, rand_seed):
        _ck(lib.ubgl_tracers_advect(self._h, sim._h, dt, rand_seed & 0xFFFFFFFF))

    def shift(self, shift):
        _ck(lib.ubgl_tracers_shift(self._h, shift))


class Items:
    """The CoItem + CoKinematicsSimple view of the reference's registry as one device array."""

    def __init__(self, items=None, device=0):
        self._h = C.c_void_p()
        _ck(lib.ubgl_items_create(device, C.byref(self._h)))
        if items is not None:
            self.set(items)

    def close(self):
        if getattr(self, "_h", None):
            lib.ubgl_items_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, items):
        items = np.ascontiguousarray(items)
        if items.dtype.itemsize != ITEM_DTYPE.itemsize:
            raise UbglError("items: need 52-byte ubgl_item records")
        self.n = len(items)
        _ck(lib.ubgl_items_upload(self._h, items.ctypes.data_as(C.c_void_p), len(items)))

    def get(self, out=None):
        """Item records back to the host; `out` (n x ITEM_DTYPE, e.g. pinned) is filled in place."""
        a = np.zeros(self.n, ITEM_DTYPE) if out is None else out
        assert a.dtype == ITEM_DTYPE and a.size >= self.n and a.flags.c_contiguous
        n = C.c_int()
        _ck(lib.ubgl_items_download(self._h, a.ctypes.data_as(C.c_void_p), self.n, C.byref(n)))
        return a

    def advect(self, sim, game_dt):
        """Simulation::advectFloatingItems (rigid bodies, advect_floating_items.cpp:16-146)."""
        _ck(lib.ubgl_items_advect(self._h, sim._h, game_dt))

    def advect_simple(self, sim, game_dt):
        """Simulation::advectFloatingItemsSimple (advect_floating_items.cpp:148-274)."""
        _ck(lib.ubgl_items_advect_simple(self._h, sim._h, game_dt))


class MG:
    """MG(width, height) -- pressure_solver.hpp:16-31, state on the GPU."""

    def __init__(self, width, height, device=0):
        self.width, self.height = width, height
        self._h = C.c_void_p()
        _ck(lib.ubgl_mg_create(width, height, device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib.ubgl_mg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, opt, val):
        _ck(lib.ubgl_mg_set_option(self._h, opt, int(val)))

    def levels(self):
        return lib.ubgl_mg_levels(self._h)

    def update_fields(self, flag):
        _ck(lib.ubgl_mg_update_fields(self._h, _fp(_f32(flag))))

    def flagc(self, level):
        w, h = C.c_int(), C.c_int()
        _ck(lib.ubgl_mg_level_size(self._h, level, C.byref(w), C.byref(h)))
        a = np.empty((h.value, w.value), np.float32)
        _ck(lib.ubgl_mg_get_flagc(self._h, level, _fp(a)))
        return a

    def solve_host(self, p, f, flag, hh, zero_gradient_bc=False):
        """MG::solve(p, f, flag, h, zeroGradientBC): p is updated in place."""
        assert p.dtype == np.float32 and p.flags["C_CONTIGUOUS"]
        _ck(lib.ubgl_mg_solve_host(self._h, _fp(p), _fp(_f32(f)), _fp(_f32(flag)), hh,
                                   int(zero_gradient_bc)))

    def set(self, p=None, f=None, flag=None):
        args = [_fp(_f32(a)) if a is not None else None for a in (p, f, flag)]
        _ck(lib.ubgl_mg_upload(self._h, *args))

    def get_p(self):
        a = np.empty((self.height, self.width), np.float32)
        _ck(lib.ubgl_mg_download_p(self._h, _fp(a)))
        return a

    def solve(self, hh, zero_gradient_bc=False, cycles=1):
        _ck(lib.ubgl_mg_solve(self._h, hh, int(zero_gradient_bc), cycles))

    def residual(self, hh):
        l2 = C.c_float()
        _ck(lib.ubgl_mg_residual(self._h, hh, C.byref(l2)))
        return l2.value

    def sync(self):
        _ck(lib.ubgl_mg_sync(self._h))

    def launch_count(self):
        return lib.ubgl_mg_launch_count(self._h)


def slab_set_row_weights(weights):
    """Relative cost of every level-0 row for the plans made from now on (None: equal rows)."""
    if weights is None:
        _ck(lib.ubgl_slab_set_row_weights(None, 0))
        return
    w = _f32(weights)
    _ck(lib.ubgl_slab_set_row_weights(_fp(w), len(w)))


def slab_plan(W, H, nranks, rank):
    """The row decomposition a (W, H, nranks) run uses -- pure host arithmetic."""
    v = [C.c_int() for _ in range(6)]
    _ck(lib.ubgl_slab_plan(W, H, nranks, rank, *[C.byref(x) for x in v]))
    k = ("dist_levels", "ghost", "own_lo", "own_hi", "st_lo", "st_hi")
    return dict(zip(k, (x.value for x in v)))


class SlabSimulation:
    """One rank's slab of a Simulation decomposed over the GPUs of a box.

    ``flag_rows`` covers this rank's STORED rows (slab_plan st_lo..st_hi).
    ``exchange`` is a callable(bytes) -> list[bytes] that all-gathers a blob
    across the ranks (torch.distributed in bench.py / the tests)."""

    def __init__(self, flag_rows, W, H, rank, nranks, exchange, pwidth=0.8, mu=0.001, device=None):
        self.W, self.H, self.rank, self.nranks = W, H, rank, nranks
        self.plan = slab_plan(W, H, nranks, rank)
        flag_rows = _f32(flag_rows)
        if flag_rows.shape != (self.plan["st_hi"] - self.plan["st_lo"], W):
            raise UbglError(f"flag rows {flag_rows.shape} do not match the stored rows of the plan")
        self.h = np.float32(pwidth) / (np.float32(W) - np.float32(1.0))
        self._h = C.c_void_p()
        _ck(lib.ubgl_slab_create(_fp(flag_rows), W, H, pwidth, mu, rank if device is None else device,
                                 rank, nranks, C.byref(self._h)))
        n = lib.ubgl_slab_ipc_size()
        blob = C.create_string_buffer(n)
        _ck(lib.ubgl_slab_ipc_export(self._h, blob))
        blobs = exchange(blob.raw)
        assert len(blobs) == nranks and all(len(b) == n for b in blobs)
        allb = C.create_string_buffer(b"".join(blobs), n * nranks)
        _ck(lib.ubgl_slab_connect(self._h, allb))

    def close(self):
        if getattr(self, "_h", None):
            lib.ubgl_slab_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def field_rows(self, field):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        _ck(lib.ubgl_slab_field_rows(self._h, field, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def set(self, field, rows):
        r0, n, w = self.field_rows(field)
        rows = _f32(rows)
        if rows.shape != (n, w):
            raise UbglError(f"field {field}: got {rows.shape}, stored rows are {(n, w)}")
        _ck(lib.ubgl_slab_upload(self._h, field, _fp(rows)))

    def set_from_global(self, field, a):
        r0, n, w = self.field_rows(field)
        self.set(field, a[r0:r0 + n])

    def get(self, field):
        """(first stored row, array of the stored rows)"""
        r0, n, w = self.field_rows(field)
        a = np.empty((n, w), np.float32)
        _ck(lib.ubgl_slab_download(self._h, field, _fp(a)))
        return r0, a

    def get_own(self, field):
        """(own_lo, rows this rank owns)"""
        r0, a = self.get(field)
        lo, hi = self.plan["own_lo"], min(self.plan["own_hi"], r0 + a.shape[0])
        return lo, a[lo - r0:hi - r0]

    def set_option(self, opt, val):
        _ck(lib.ubgl_slab_set_option(self._h, opt, int(val)))

    def set_sinks(self, xyz):
        xyz = _f32(np.asarray(xyz, np.float32).reshape(-1, 3))
        _ck(lib.ubgl_slab_set_sinks(self._h, _fp(xyz), len(xyz)))

    def step(self, dt):
        _ck(lib.ubgl_slab_step(self._h, dt))

    def step_host(self, dt, vx_accum=None, vy_accum=None, vx=None, vy=None, p=None, vx_current=None,
                  vy_current=None):
        """One step as a host caller sees it; every array covers this rank's STORED rows of its field
        (field_rows), inputs are read whole, outputs are written on the own rows."""
        m = HostMirrors()
        ids = dict(vx_accum=VX_ACCUM, vy_accum=VY_ACCUM, vx=VX, vy=VY, p=P, vx_current=VX_CURRENT,
                   vy_current=VY_CURRENT)
        for name, a in (("vx_accum", vx_accum), ("vy_accum", vy_accum), ("vx", vx), ("vy", vy), ("p", p),
                        ("vx_current", vx_current), ("vy_current", vy_current)):
            if a is not None:
                r0, n, w = self.field_rows(ids[name])
                if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"] or a.shape != (n, w):
                    raise UbglError(f"{name}: need a C-contiguous float32 array of the stored rows {(n, w)}")
            setattr(m, name, _fp(a) if a is not None else None)
        m.flag = None
        _ck(lib.ubgl_slab_step_host(self._h, dt, C.byref(m)))

    def sync(self):
        _ck(lib.ubgl_slab_sync(self._h))

    def residual_sumsq(self):
        v = C.c_double()
        _ck(lib.ubgl_slab_residual_sumsq(self._h, C.byref(v)))
        return v.value

    def launch_count(self):
        return lib.ubgl_slab_launch_count(self._h)

    def stream(self):
        return lib.ubgl_slab_stream(self._h)

    def stats(self):
        a, b = C.c_longlong(), C.c_longlong()
        _ck(lib.ubgl_slab_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def profile(self, on):
        _ck(lib.ubgl_slab_profile(self._h, int(on)))

    def kernel_stats(self):
        """{(kind_name, level): (launches, total_ms)} of this rank's profiled region."""
        out = {}
        for k in range(lib.ubgl_num_kernel_kinds()):
            for l in range(16):
                n, ms = C.c_longlong(), C.c_double()
                _ck(lib.ubgl_slab_kernel_stats(self._h, k, l, C.byref(n), C.byref(ms)))
                if n.value:
                    out[(lib.ubgl_kernel_kind_name(k).decode(), l)] = (n.value, ms.value)
        return out


# ---- pressure_solver.cpp free functions -------------------------------------
def rbgs(p, f, flag, hh, alpha=1.0, sweeps=1):
    p = _f32(p).copy()
    H, W = p.shape
    _ck(lib.ubgl_rbgs(_fp(p), _fp(_f32(f)), _fp(_f32(flag)), W, H, hh, alpha, sweeps))
    return p


def residual(p, f, flag, hh):
    p = _f32(p)
    H, W = p.shape
    r = np.zeros_like(p)
    l2 = C.c_float()
    _ck(lib.ubgl_residual(_fp(p), _fp(_f32(f)), _fp(_f32(flag)), _fp(r), W, H, hh, C.byref(l2)))
    return r, l2.value


def restrict(r):
    r = _f32(r)
    H, W = r.shape
    rc = np.zeros((H // 2, W // 2), np.float32)
    _ck(lib.ubgl_restrict(_fp(r), W, H, _fp(rc)))
    return rc


def prolongate(ec, flagc, flag):
    flag = _f32(flag)
    H, W = flag.shape
    e = np.full((H, W), 7.0, np.float32)
    _ck(lib.ubgl_prolongate(_fp(e), W, H, _fp(_f32(ec)), _fp(_f32(flagc)), _fp(flag)))
    return e


def correct(p, e):
    p = _f32(p).copy()
    H, W = p.shape
    _ck(lib.ubgl_correct(_fp(p), _fp(_f32(e)), W, H))
    return p


def zero_gradient_bc(p):
    p = _f32(p).copy()
    H, W = p.shape
    _ck(lib.ubgl_zero_gradient_bc(_fp(p), W, H))
    return p
