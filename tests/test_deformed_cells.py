"""DeformedCells_cli (src/bin/DeformedCells_cli.cc:136-418 of the reference).
CPU: the oracle's two routes to the homogenized tensor of a linearly deformed cell -- solving on the deformed
geometry, and solving on the undeformed cell with the pulled-back material followed by a push-forward -- are the
same discrete problem and must agree to rounding (theory-pinned KAT); tiling, usage errors.  The CLI against the
oracle on the GPU: tests/test_zz_deformed_cells_gpu.py."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import meshfem_oracle as orc
from util import B9CREATOR, ROOT

BIN = os.path.join(ROOT, "bin")
J2 = np.array([[1.2, 0.3], [0.1, 0.9]])
J3 = np.array([[1.1, 0.2, 0.0], [0.05, 0.9, 0.1], [0.0, 0.15, 1.05]])


def _run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, **kw)


def _cell(lib_built, N, n, hole):
    from meshfem_b200 import hostlib
    return hostlib.perforated_cell(N, n, hole)


@pytest.mark.parametrize("N,n,hole,deg", [(2, 6, 2, 1), (2, 6, 2, 2), (3, 4, 2, 1)])
def test_oracle_deformed_cell_direct_equals_transform_version(lib_built, N, n, hole, deg):
    V, T = _cell(lib_built, N, n, hole).arrays()
    D = orc.isotropic_D(N, 200.0, 0.35)
    J = J2 if N == 2 else J3
    E_direct, w_d, _ = orc.deformed_cell_homogenization(N, deg, V, T, D, J, transform_version=False)
    E_pull, w_p, _ = orc.deformed_cell_homogenization(N, deg, V, T, D, J, transform_version=True)
    assert np.abs(E_direct - E_pull).max() <= 1e-12 * np.abs(E_direct).max()
    assert np.allclose(E_direct, E_direct.T, atol=1e-10 * np.abs(E_direct).max())
    assert np.linalg.eigvalsh(E_direct).min() > 0
    # identity deformation = the plain homogenization
    sim = orc.Simulator(N, deg, V, T); sim.set_material(D)
    E0 = orc.homogenized_tensor_displacement_form(sim, orc.solve_cell_problems(sim))
    E_id, _, _ = orc.deformed_cell_homogenization(N, deg, V, T, D, np.eye(N))
    assert np.abs(E_id - E0).max() <= 1e-12 * np.abs(E0).max()
    # a rotation of the cell rotates the tensor
    th = 0.4
    Q = np.eye(N); Q[0, 0] = Q[1, 1] = np.cos(th); Q[0, 1] = -np.sin(th); Q[1, 0] = np.sin(th)
    E_rot, _, _ = orc.deformed_cell_homogenization(N, deg, V, T, D, Q)
    assert np.abs(E_rot - orc.transform_tensor(N, E0, Q)).max() <= 1e-11 * np.abs(E0).max()


@pytest.mark.parametrize("N,n,hole,deg", [(2, 6, 2, 2), (3, 4, 2, 1), (3, 4, 2, 2)])
def test_host_deformed_displacement_form_matches_oracle(lib_built, N, n, hole, deg):
    """The C++ host half of --homogenize (periodic matching on the undeformed cell, updateMeshNodePositions, boundary
    normals / volumes of the deformed geometry, Eh over |bbox| det J) fed with the oracle's fluctuation displacements."""
    raw = _cell(lib_built, N, n, hole)
    V, T = raw.arrays()
    D = orc.isotropic_D(N, 200.0, 0.35)
    J = J2 if N == 2 else J3
    E_ref, w, sim = orc.deformed_cell_homogenization(N, deg, V, T, D, J)
    Eh, nodes = raw.deformed_displacement_form(deg, D, J, np.array(w))
    assert np.abs(nodes - sim.mesh.nodes).max() < 1e-14
    assert np.abs(Eh - E_ref).max() <= 1e-13 * np.abs(E_ref).max()


def test_transform_tensor_properties():
    D = orc.material_from_json(3, {"type": "orthotropic", "young": [200, 120, 80], "poisson": [0.3, 0.2, 0.12, 0.3, 0.3, 0.18],
                                   "shear": [45, 35, 60]})
    A, B = J3, J3.T @ J3 + np.eye(3)
    assert np.allclose(orc.transform_tensor(3, orc.transform_tensor(3, D, A), B), orc.transform_tensor(3, D, B @ A), rtol=1e-12)
    assert np.allclose(orc.transform_tensor(3, orc.transform_tensor(3, D, A), np.linalg.inv(A)), D, rtol=1e-11, atol=1e-11)
    iso = orc.isotropic_D(3, 200.0, 0.35)
    th = 0.7
    Q = np.array([[1, 0, 0], [0, np.cos(th), -np.sin(th)], [0, np.sin(th), np.cos(th)]])
    assert np.allclose(orc.transform_tensor(3, iso, Q), iso, atol=1e-12)
    assert np.isclose(orc.frobenius_norm_sq(3, orc.transform_tensor(3, D, Q)), orc.frobenius_norm_sq(3, D), rtol=1e-13)


def test_deformed_cells_cli_tiling_and_usage(lib_built, tmp_path):
    mesh = str(tmp_path / "sq.msh")
    assert _run([os.path.join(BIN, "grid"), "2x2", "-t", mesh]).returncode == 0
    out = str(tmp_path / "tiled.msh")
    r = _run([os.path.join(BIN, "DeformedCells_cli"), mesh, "-j", "1 0.5 0 1", "-t", "2 3", "-o", out])
    assert r.returncode == 0, r.stderr
    V0, T0, _ = orc.read_msh(mesh)
    V, T, _ = orc.read_msh(out)
    nv, ne = V0.shape[0], T0.shape[0]
    assert V.shape[0] == 6 * nv and T.shape[0] == 6 * ne      # copies are not glued (reference: "TODO: merge")
    J = np.array([[1.0, 0.5], [0.0, 1.0]])
    center = 0.5 * (V0[:, :2].min(0) + V0[:, :2].max(0))
    dims = V0[:, :2].max(0) - V0[:, :2].min(0)
    copy = 0
    for i in range(2):
        for j in range(3):
            expect = (V0[:, :2] - center) @ J.T + J @ (np.array([i, j]) * dims)
            assert np.allclose(V[copy * nv:(copy + 1) * nv, :2], expect, atol=1e-14)
            assert np.array_equal(T[copy * ne:(copy + 1) * ne], T0 + copy * nv)
            copy += 1
    # usage errors (DeformedCells_cli.cc:78-101)
    r = _run([os.path.join(BIN, "DeformedCells_cli"), mesh, "-t", "2 2"])
    assert r.returncode == 1 and "no operation requested" in r.stderr and "must specify either deformation jacobian" in r.stderr
    r = _run([os.path.join(BIN, "DeformedCells_cli"), mesh, "-j", "1 0 0 1", "-t", "2 2", "--homogenize"])
    assert r.returncode == 1 and "do not specify both tiling and homogenization" in r.stderr
    r = _run([os.path.join(BIN, "DeformedCells_cli"), mesh, "-j", "1 0 0", "-t", "2 2", "-o", out])
    assert r.returncode == 1 and "Invalid deformation jacobian" in r.stderr
    r = _run([os.path.join(BIN, "DeformedCells_cli"), mesh, "-j", "1 0 0 1", "-t", "2 0", "-o", out])
    assert r.returncode == 1 and "Invalid number of tilings" in r.stderr
