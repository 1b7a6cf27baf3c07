"""ubootgl_b200 -- B200-native (sm_100a) fluid-step hot path of te42kyfo/ubootgl.

The product is ``_lib/libubgl.so`` (hand-written CUDA behind the C ABI in
``include/ubgl.h``) plus the C++ drop-in mirror of the reference's
``Simulation`` / ``MG`` classes in ``host/``.  This Python package is only the
ctypes face of that C ABI used by tests/ and bench.py; it contains no compute
and no CPU fallback: importing ``capi`` without the built library raises.
"""
from . import capi  # noqa: F401
from .capi import MG, DisplayArray, Items, Simulation, SlabSimulation, Tracers, UbglError, lib, slab_plan, slab_set_row_weights  # noqa: F401
