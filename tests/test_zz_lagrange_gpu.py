"""GPU: systems with Lagrange-multiplier rows through the reference-facing surface (Simulate_cli with
`no_rigid_motion`).  The host resolves the rows around the device PCG
(include/MeshFEM/RigidMotionConstraints.hh); the oracle solves the reference's saddle-point matrix directly.
Tolerance 1e-6 relative L2 on the displacements (north-star gate)."""
import json
import os
import subprocess

import numpy as np
import pytest

import meshfem_oracle as orc
from util import B9CREATOR, ROOT, read_msh_fields, rel_l2

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "bin")


def _run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600)


def _write(path, obj):
    with open(path, "w") as f:
        json.dump(obj, f)
    return str(path)


def _oracle(N, deg, sizes, bc):
    V, T = orc.grid_simplices(list(sizes))
    return orc.simulate(N, deg, V, T, orc.material_from_json(N, B9CREATOR), bc)


@pytest.mark.parametrize("deg", [1, 2])
def test_simulate_cli_no_rigid_motion_3d(lib_built, tmp_path, deg):
    """A floating bar: unbalanced force on one face, six Lagrange rows (three rotations, three translations)."""
    bc = {"no_rigid_motion": True, "regions": [
        {"type": "force", "value": [3.0, 1.0, 0.0], "box%": {"minCorner": [0.999, -0.01, -0.01], "maxCorner": [1.001, 1.01, 1.01]}},
        {"type": "traction", "value": [0.0, 0.0, 0.2], "box%": {"minCorner": [-0.001, -0.01, -0.01], "maxCorner": [0.001, 1.01, 1.01]}}]}
    mesh = str(tmp_path / "bar.msh")
    assert _run([os.path.join(BIN, "grid"), "6x2x2", "-t", mesh]).returncode == 0
    out = str(tmp_path / "out.msh")
    r = _run([os.path.join(BIN, "Simulate_cli"), mesh, "-m", _write(tmp_path / "m.material", B9CREATOR), "-b", _write(tmp_path / "c.bc", bc),
              "-d", str(deg), "-o", out, "-D"])
    assert r.returncode == 0, r.stderr + r.stdout
    f = read_msh_fields(out)
    ref = _oracle(3, deg, (6, 2, 2), bc)
    assert rel_l2(f["u"], ref["u"]) < 1e-6
    sim = ref["sim"]
    _, _, C, d = sim.constraints()
    assert np.abs(C @ f["u"].reshape(-1)).max() < 1e-8 * np.abs(f["u"]).max() * np.abs(C).max() * C.shape[1]


def test_simulate_cli_unconstrained_translation_2d(lib_built, tmp_path):
    """y fixed on the bottom edge, x translation free and no pin: one translation row."""
    bc = {"regions": [{"type": "dirichlety", "value": [0, "0.01*x", 0], "box%": {"minCorner": [-0.01, -0.001], "maxCorner": [1.01, 0.001]}},
                      {"type": "force", "value": [1.0, -2.0, 0], "box%": {"minCorner": [-0.01, 0.999], "maxCorner": [1.01, 1.001]}}]}
    mesh = str(tmp_path / "sq.msh")
    assert _run([os.path.join(BIN, "grid"), "8x5", "-t", mesh]).returncode == 0
    out = str(tmp_path / "out.msh")
    r = _run([os.path.join(BIN, "Simulate_cli"), mesh, "-m", _write(tmp_path / "m.material", B9CREATOR), "-b", _write(tmp_path / "c.bc", bc),
              "-d", "2", "-o", out, "-D"])
    assert r.returncode == 0, r.stderr + r.stdout
    f = read_msh_fields(out)
    ref = _oracle(2, 2, (8, 5), bc)
    assert rel_l2(f["u"][:, :2], ref["u"]) < 1e-6
