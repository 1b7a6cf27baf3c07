"""The C++ drop-in mirror of the reference's Simulation / MG classes
(ubootgl_b200/host/): builds against include/ubgl.h, and -- on the GPU --
reproduces the reference's mgtest known answers and a game-like call sequence
(step, accumulator scatter, explosion sink, terrain edit via setGrids +
mg.updateFields) checked against the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "ubootgl_b200", "host")
BUILD = os.path.join(HOST, "_build")


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "ubootgl_b200")])
    subprocess.check_call(["make", "-s", "-C", HOST])


def test_host_mirror_builds_and_keeps_the_reference_api():
    _build()
    syms = subprocess.run(["nm", "-C", os.path.join(BUILD, "libubgl_host.a")], capture_output=True,
                          text=True, check=True).stdout
    for name in ("Simulation::step(float)", "Simulation::setVBCs()", "Simulation::setPBC()",
                 "Simulation::saveCurrentVelocityFields()", "Simulation::psampleFlagLinear(",
                 "Simulation::psampleFlagNormal(", "Simulation::psampleFlagNearest(",
                 "Simulation::setGrids(", "Simulation::diffuse()", "Simulation::advect()",
                 "Simulation::project()", "Simulation::applyAccumulatedVelocity()",
                 "MG::MG(int, int", "MG::updateFields(Single2DGrid&)",
                 "MG::solve(Single2DGrid&, Single2DGrid&, Single2DGrid&, float, bool)",
                 "rbgs(Single2DGrid&, Single2DGrid&, Single2DGrid&, float, float)",
                 "calculateResidualField(Single2DGrid&, Single2DGrid&, Single2DGrid&, Single2DGrid&, float)"):
        assert name in syms, name
    for exe in ("mgtest", "simdemo"):
        assert os.access(os.path.join(BUILD, exe), os.X_OK)


def demo_flag(W, H):
    flag = np.ones((H, W), np.float32)
    flag[0, :] = 0
    flag[H - 1, :] = 0
    flag[H // 4:H // 2, W // 5:W // 5 + W // 10] = 0
    flag[H // 2:3 * H // 4, W // 2:W // 2 + W // 12] = 0
    return flag


@pytest.mark.gpu
def test_mgtest_known_answers(ubgl):
    """mgtest.cpp:34-53 through the C++ drop-in: residual history within 1 %."""
    _build()
    out = subprocess.run([os.path.join(BUILD, "mgtest"), "1025", "--resident"], capture_output=True,
                         text=True, check=True, timeout=300).stdout.split("\n")
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "mgtest_kat.json")))
    hist = [float(out[0].split(":")[1])] + [float(l) for l in out[1:6]]
    for a, b in zip(hist, kat["residual_history"]):
        assert abs(a - b) <= 0.01 * b, (hist, kat["residual_history"])
    assert abs(float(out[6]) - kat["scaled_error"]) <= 0.01 * kat["scaled_error"]


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,steps", [(160, 96, 4), (258, 131, 5)])
def test_simdemo_matches_oracle(ubgl, port, tmp_path, W, H, steps):
    from oracle import bind as ob
    _build()
    prefix = str(tmp_path / "demo")
    subprocess.run([os.path.join(BUILD, "simdemo"), str(W), str(H), str(steps), prefix], check=True,
                   timeout=300)
    flag = demo_flag(W, H)
    O = port.Sim(flag, 0.8, 0.001)
    for s in range(steps):
        ax = O.get(ob.VX_ACCUM)
        ay = O.get(ob.VY_ACCUM)
        ax[H // 3, W // 3] += np.float32(0.02)
        ay[H // 3, W // 3] -= np.float32(0.01)
        O.set(ob.VX_ACCUM, ax)
        O.set(ob.VY_ACCUM, ay)
        if s == 1:
            O.add_sink(0.5 * 0.8, np.float32(0.5) * np.float32(0.8) * H / W, 120.0)
        if s == 2:
            # Simulation::setGrids (simulation.hpp:82-98) + mg.updateFields
            fl, vx, vy, p = O.get(ob.FLAG), O.get(ob.VX), O.get(ob.VY), O.get(ob.P)
            fl[H // 4:H // 4 + 4, W // 5:W // 5 + 4] = 1
            for y in range(3 * H // 5, 3 * H // 5 + 3):
                for x in range(3 * W // 4, 3 * W // 4 + 3):
                    fl[y, x] = 0
                    vx[y, x] = 0
                    vx[y, x - 1] = 0
                    vy[y, x] = 0
                    vy[y - 1, x] = 0
                    p[y, x] = 0
            O.set(ob.VX, vx)
            O.set(ob.VY, vy)
            O.set(ob.P, p)
            O.update_flag(fl)
        O.step(0.001)
    got = {
        "vx": np.fromfile(prefix + ".vx", np.float32).reshape(H, W - 1),
        "vy": np.fromfile(prefix + ".vy", np.float32).reshape(H - 1, W),
        "p": np.fromfile(prefix + ".p", np.float32).reshape(H, W),
        "vxc": np.fromfile(prefix + ".vxc", np.float32).reshape(H, W - 1),
        "flag": np.fromfile(prefix + ".flag", np.float32).reshape(H, W),
    }
    assert (got["flag"] == O.get(ob.FLAG)).all()
    assert cases.rel_l2(got["vx"], O.get(ob.VX)) <= 2e-5
    assert cases.rel_l2(got["vy"], O.get(ob.VY)) <= 2e-5 * max(1.0, np.linalg.norm(O.get(ob.VX)) / max(np.linalg.norm(O.get(ob.VY)), 1e-30))
    assert cases.rel_l2(got["p"], O.get(ob.P)) <= 5e-5
    assert cases.rel_l2(got["vxc"], O.get(ob.VX_CURRENT)) <= 2e-5
