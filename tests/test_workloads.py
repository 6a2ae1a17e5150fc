"""tools/workloads.py (the vectorised cantilever boundary conditions that feed every bench number) against the
oracle's general implementation of the reference semantics: .bc parsing (BoundaryConditions.cc:217-388), Dirichlet
nodes by position / Neumann elements by vertex barycentre / `force` divided by the region area
(LinearElasticity.hh:881-1027), m_getDirichletVarsAndValues (:1469-1518) and neumannLoad (:703-717, 341-347).
Also pins the materials of BASELINE.md against the oracle's .material path."""
import os
import sys

import numpy as np
import pytest

import meshfem_oracle as orc
from util import ORTHO, ROOT, cantilever_problem

sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("deg", [1, 2])
@pytest.mark.parametrize("sizes", [(6, 3, 2), (10, 2, 2), (5, 5, 5)])
def test_cantilever_inputs_match_oracle_boundary_conditions(lib_built, deg, sizes):
    import workloads as wl
    m = wl.grid_femmesh(sizes, deg)
    fixed, vals, f = wl.cantilever_inputs(m)
    sim, ofixed, ovals, of = cantilever_problem(3, deg, sizes)
    # same mesh, same node numbering (FEMMesh.inl:17-37), so the per-node arrays are directly comparable
    assert m.num_nodes == sim.mesh.num_nodes and np.array_equal(np.asarray(m.elem_nodes), np.asarray(sim.mesh.elem_nodes))
    assert np.allclose(m.nodes, sim.mesh.nodes, rtol=0, atol=0)
    o = np.argsort(fixed); oo = np.argsort(ofixed)
    assert np.array_equal(np.asarray(fixed)[o], np.asarray(ofixed)[oo])
    assert np.array_equal(np.asarray(vals)[o], np.asarray(ovals)[oo])
    of = np.asarray(of).reshape(-1, 3)
    assert np.abs(f - of).max() <= 1e-15 * np.abs(of).max()
    # total force = the .bc's `force` value; quadratic faces load the edge nodes only (Functions.hh:257-274)
    assert np.allclose(f.sum(axis=0), [0.0, -10.0, 0.0], atol=1e-12)
    if deg == 2:
        nv = int(np.asarray(m.elem_nodes)[:, :4].max()) + 1
        assert np.abs(f[:nv]).max() == 0.0


def test_bench_materials_match_oracle_material_files():
    import workloads as wl
    iso = orc.material_from_json(3, {"type": "isotropic_material", "dim": 3, "young": 200.0, "poisson": 0.35})
    assert np.abs(wl.material("iso") - iso).max() <= 1e-12 * np.abs(iso).max()
    ortho = orc.material_from_json(3, ORTHO)
    assert np.abs(wl.material("ortho") - ortho).max() <= 1e-12 * np.abs(ortho).max()
    assert np.linalg.eigvalsh(wl.material("ortho")).min() > 0


def test_configs_are_the_baseline_sizes():
    import workloads as wl
    sizes = {k: 24 * g[0] * g[1] * g[2] for k, (g, _, _) in wl.CONFIGS.items()}
    assert sizes == {"cfg2": 960000, "cfg3": 2109120, "cfg5": 10222080}
    assert [wl.CONFIGS[k][1] for k in ("cfg2", "cfg3", "cfg5")] == [1, 2, 2]
