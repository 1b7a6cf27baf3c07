"""GPU: the row-pair advect kernels (default) must reproduce the one-face-per-thread kernels
BIT FOR BIT -- same per-face arithmetic, only the tap loads are shared between the two rows of
a pair.  The variant is fixed per process (UBGL_ADVECT_VARIANT), hence two subprocesses."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pair_kernels_equal_single_face_kernels(ubgl, tmp_path):
    outs = []
    for variant in ("1", "2"):
        out = str(tmp_path / f"advect_v{variant}.npz")
        env = dict(os.environ, UBGL_ADVECT_VARIANT=variant)
        subprocess.run([sys.executable, os.path.join(ROOT, "tests", "advect_dump.py"), out], check=True, env=env,
                       timeout=300)
        outs.append(np.load(out))
    a, b = outs
    assert sorted(a.files) == sorted(b.files) and len(a.files) > 0
    for k in a.files:
        x, y = a[k], b[k]
        same = (x.view(np.uint32) == y.view(np.uint32)) | ((x == 0) & (y == 0))
        assert same.all(), (k, float(np.abs(x - y).max()))


def test_divergence_in_the_advect_epilogue_equals_the_separate_pass(ubgl, tmp_path):
    """UBGL_ADVECT_DIV=1: k_advect_xy<DIV> writes f = -(1/h) div v of its cells from the freshly advected faces
    (left lane by shuffle, lower row through shared memory), k_divergence_edges the CTA edges and the ring next to
    the BC faces after setVBCs -- bit for bit the f, p and velocities of the default (k_divergence4 as its own pass)."""
    outs = []
    for div in ("0", "1"):
        out = str(tmp_path / f"advect_div{div}.npz")
        env = dict(os.environ, UBGL_ADVECT_DIV=div)
        env.pop("UBGL_ADVECT_VARIANT", None)
        subprocess.run([sys.executable, os.path.join(ROOT, "tests", "advect_dump.py"), out], check=True, env=env,
                       timeout=300)
        outs.append(np.load(out))
    a, b = outs
    assert sorted(a.files) == sorted(b.files) and any(k.endswith("_f2") for k in a.files)
    for k in a.files:
        x, y = a[k], b[k]
        same = (x.view(np.uint32) == y.view(np.uint32)) | ((x == 0) & (y == 0))
        assert same.all(), (k, float(np.abs(x - y).max()))


def test_programmatic_dependent_launch_changes_nothing(ubgl, tmp_path):
    """UBGL_PDL=0 (ordinary launches) vs the default (the multigrid passes, border kernels and sink stamps launched
    with programmatic stream serialization, parked at griddepcontrol.wait): the same bits after the advect stage
    and two whole steps at six sizes and two CFLs, graph replay included on the small grids."""
    outs = []
    for pdl in ("0", "1"):
        out = str(tmp_path / f"pdl{pdl}.npz")
        env = dict(os.environ, UBGL_PDL=pdl)
        env.pop("UBGL_ADVECT_VARIANT", None)
        subprocess.run([sys.executable, os.path.join(ROOT, "tests", "advect_dump.py"), out], check=True, env=env,
                       timeout=300)
        outs.append(np.load(out))
    a, b = outs
    assert sorted(a.files) == sorted(b.files) and len(a.files) > 0
    for k in a.files:
        assert (a[k].view(np.uint32) == b[k].view(np.uint32)).all(), k
