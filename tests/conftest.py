import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    """The plain-C restatement (oracle/ubgl_oracle.c), built on demand."""
    from oracle import bind
    if not bind.have_port():
        bind.build(ref=False)
    return bind.Port()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference TUs (oracle/_ref); skipped when not built/shipped."""
    from oracle import bind
    if not bind.have_ref():
        if os.path.exists("/root/reference/simulation.cpp"):
            bind.build(ref=True)
        else:
            pytest.skip("oracle/_ref not available on this box")
    return bind.Ref()


@pytest.fixture(scope="session")
def ubgl():
    """The product: ctypes face of libubgl.so.  GPU tests call through this."""
    import ubootgl_b200
    if ubootgl_b200.lib.ubgl_device_count() < 1:
        pytest.fail("GPU test selected but libubgl sees no CUDA device")
    return ubootgl_b200


@pytest.fixture(scope="session")
def ref_strict():
    """The unmodified reference fluid TUs compiled IEEE-strict (-O2 -ffp-contract=off) instead of
    -Ofast: the yardstick for the reference's own rounding sensitivity."""
    from oracle import bind
    if not bind.have_ref_strict():
        if os.path.exists("/root/reference/simulation.cpp"):
            bind.build(ref=True)
        else:
            pytest.skip("oracle/_ref/libubgl_ref_strict.so not available on this box")
    return bind.RefStrict()


@pytest.fixture(scope="session")
def glsl():
    """The reference's unmodified GLSL compute shaders compiled as C++ (oracle/_ref/libubgl_glsl.so)."""
    from oracle import bind
    if not bind.have_glsl():
        if os.path.exists("/root/reference/advect_tracer_points.cs"):
            bind.build(ref=True)
        if not bind.have_glsl():
            pytest.skip("oracle/_ref/libubgl_glsl.so not available on this box")
    return bind.Glsl()
