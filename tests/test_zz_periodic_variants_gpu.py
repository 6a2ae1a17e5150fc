"""GPU: PeriodicHomogenization_cli with --ignorePeriodicMismatch, --manualPeriodicVertices and --m2mstress against
the oracle (CPU side of the same features: tests/test_periodic_variants.py)."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import meshfem_oracle as orc
from util import B9CREATOR, ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "bin")


def _run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, **kw)


def _eh(stdout, F):
    lines = stdout.split("\n")
    k = lines.index("Homogenized elasticity tensor:")
    return np.array([[float(x) for x in lines[k + 1 + i].split()] for i in range(F)])


@pytest.fixture(scope="module")
def cell(lib_built, tmp_path_factory):
    from meshfem_b200 import hostlib
    d = tmp_path_factory.mktemp("cell")
    raw = hostlib.perforated_cell(3, 4, 2)
    mesh = str(d / "cell.msh")
    raw.save(mesh)
    mat = str(d / "m.material")
    with open(mat, "w") as f:
        json.dump(B9CREATOR, f)
    V, T = raw.arrays()
    sim = orc.Simulator(3, 1, V, T)
    sim.set_material(orc.material_from_json(3, B9CREATOR))
    w = orc.solve_cell_problems(sim)
    Eh = orc.homogenized_tensor_displacement_form(sim, w)
    return dict(dir=d, mesh=mesh, mat=mat, sim=sim, w=w, Eh=Eh)


def test_ignore_periodic_mismatch_on_matching_cell(cell):
    r = _run([os.path.join(BIN, "PeriodicHomogenization_cli"), cell["mesh"], "-m", cell["mat"], "-d", "1", "--ignorePeriodicMismatch"])
    assert r.returncode == 0, r.stderr + r.stdout
    assert np.abs(_eh(r.stdout, 6) - cell["Eh"]).max() < 1e-7 * np.abs(cell["Eh"]).max()


def test_manual_periodic_vertices(cell):
    sim = cell["sim"]
    dof = sim.dof_for_node
    pairs = []
    for d in np.unique(dof):
        nodes = np.nonzero(dof == d)[0]
        pairs += [(int(nodes[k]), int(nodes[k + 1])) for k in range(nodes.size - 1)]
    path = str(cell["dir"] / "pairs.txt")
    with open(path, "w") as f:
        f.write("".join(f"{a} {b}\n" for a, b in pairs))
    r = _run([os.path.join(BIN, "PeriodicHomogenization_cli"), cell["mesh"], "-m", cell["mat"], "-d", "1", "--manualPeriodicVertices", path])
    assert r.returncode == 0, r.stderr + r.stdout
    assert "temporary hack" in r.stderr
    assert np.abs(_eh(r.stdout, 6) - cell["Eh"]).max() < 1e-7 * np.abs(cell["Eh"]).max()


def test_m2mstress(cell):
    out = str(cell["dir"] / "m2m.txt")
    r = _run([os.path.join(BIN, "PeriodicHomogenization_cli"), cell["mesh"], "-m", cell["mat"], "-d", "1", "-M", out], cwd=str(cell["dir"]))
    assert r.returncode == 0, r.stderr + r.stdout
    sim = cell["sim"]
    M = orc.macro_to_micro_stress_tensors(sim, cell["w"], cell["Eh"])
    G = orc.macro_to_micro_strain_tensors(sim, cell["w"])
    for path, ref in ((out, M), (str(cell["dir"] / "gtensors.txt"), G)):
        lines = [ln for ln in open(path).read().splitlines() if ln.strip()]
        assert len(lines) == sim.mesh.num_elements
        for e in (0, len(lines) // 3, len(lines) - 1):
            vals = np.array([float(x) for x in re.findall(r"[-+0-9.eE]+", lines[e])]).reshape(3, 3, 3, 3)
            want = orc.unflatten_rank4(3, ref[e])
            assert np.abs(vals - want).max() < 1e-6 * np.abs(want).max()
