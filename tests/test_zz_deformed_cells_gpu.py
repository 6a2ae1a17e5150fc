"""GPU: DeformedCells_cli --homogenize (direct, --transformVersion, -p) against the oracle
(oracle: deformed_cell_homogenization; CPU side of the same feature: tests/test_deformed_cells.py).
Tolerance 1e-7 relative on the homogenized tensor."""
import json
import os
import re

import numpy as np
import pytest

import meshfem_oracle as orc
from test_deformed_cells import BIN, J2, J3, _cell, _run
from util import B9CREATOR

pytestmark = pytest.mark.gpu


def _parse_tensor(stdout, F):
    lines = stdout.split("Elasticity tensor:")[1].strip().splitlines()[:F]
    return np.array([[float(x) for x in ln.split()] for ln in lines])


@pytest.mark.parametrize("N,n,hole,deg", [(2, 8, 4, 2), (3, 4, 2, 2), (3, 4, 2, 1)])
def test_deformed_cells_cli_homogenize_matches_oracle(lib_built, tmp_path, N, n, hole, deg):
    raw = _cell(lib_built, N, n, hole)
    mesh = str(tmp_path / "cell.msh")
    raw.save(mesh)
    V, T = raw.arrays()
    J = J2 if N == 2 else J3
    D = orc.material_from_json(N, B9CREATOR)
    mat = str(tmp_path / "m.material")
    with open(mat, "w") as f:
        json.dump(B9CREATOR, f)
    E_ref, w_ref, _ = orc.deformed_cell_homogenization(N, deg, V, T, D, J)
    jac = " ".join(repr(float(x)) for x in J.reshape(-1))
    F = N * (N + 1) // 2
    dump = str(tmp_path / "eh.json")
    out = str(tmp_path / "fields.msh")
    r = _run([os.path.join(BIN, "DeformedCells_cli"), mesh, "-m", mat, "-j", jac, "-d", str(deg), "--homogenize", "-o", out, "--dumpJson", dump])
    assert r.returncode == 0, r.stderr + r.stdout
    E_direct = _parse_tensor(r.stdout, F)
    assert np.abs(E_direct - E_ref).max() <= 1e-7 * np.abs(E_ref).max()
    data = json.load(open(dump))
    assert np.allclose(np.array(data["elasticity_tensor"]).reshape(F, F), E_direct, rtol=1e-12)
    assert len(data["homogenized_moduli"]) == (4 if N == 2 else 9)
    moduli = [float(x) for x in re.search(r"Homogenized Moduli: (.*)", r.stdout).group(1).split()]
    assert np.allclose(moduli, data["homogenized_moduli"], rtol=1e-12)
    # the fluctuation fields were written on the DEFORMED geometry
    from util import read_msh_fields
    f = read_msh_fields(out)
    assert "w_ij0" in f and "load_ij 0" in f and "strain w_ij 0" in f
    # transform version: same tensor through the pulled-back material
    r2 = _run([os.path.join(BIN, "DeformedCells_cli"), mesh, "-m", mat, "-j", jac, "-d", str(deg), "--homogenize", "--transformVersion"])
    assert r2.returncode == 0, r2.stderr + r2.stdout
    assert np.abs(_parse_tensor(r2.stdout, F) - E_ref).max() <= 1e-7 * np.abs(E_ref).max()


def test_deformed_cells_cli_parametrized_transform(lib_built, tmp_path):
    raw = _cell(lib_built, 2, 8, 4)
    mesh = str(tmp_path / "cell.msh")
    raw.save(mesh)
    V, T = raw.arrays()
    D = orc.isotropic_D(2, 1.0, 0.3)                      # default material
    r = _run([os.path.join(BIN, "DeformedCells_cli"), mesh, "-p", "--homogenize", "--transformVersion", "-d", "1"],
             input="# theta lambda\n0.3 1.25\n\n1.0 0.8\n")
    assert r.returncode == 0, r.stderr + r.stdout
    rows = [ln.split("\t") for ln in r.stdout.splitlines() if ln.count("\t") >= 13]
    assert len(rows) == 2
    for row, (theta, lam) in zip(rows, [(0.3, 1.25), (1.0, 0.8)]):
        c, s = np.cos(theta), np.sin(theta)
        R = np.array([[c, -s], [s, c]])
        J = R @ np.diag([lam, 1.0]) @ R.T
        E_ref, _, _ = orc.deformed_cell_homogenization(2, 1, V, T, D, J, transform_version=True)
        vals = [float(x) for x in row]
        assert abs(vals[0] - theta) < 1e-12 and abs(vals[1] - lam) < 1e-12
        upper = [E_ref[i, j] for i in range(3) for j in range(i, 3)]
        assert np.allclose(vals[2:8], upper, rtol=1e-7, atol=1e-9)
