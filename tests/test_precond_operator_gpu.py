"""GPU: the preconditioner INSIDE the device PCG (block-Jacobi + level-1 small boxes + dense level of large boxes;
csrc/coarse.inl) against its numpy restatement (tools/emulate_multilevel.py) as an operator: z = M^-1 r and r.z for
random r through mfem_b200_apply_preconditioner, which runs the PCG's own start-up kernels (fused update+restriction,
level-1 kernel, dense GEMV, fused direction+prolongation).  Also: the operator is symmetric, and r.z is r.(M^-1 r)."""
import os
import sys

import numpy as np
import pytest

from util import ROOT, cantilever_problem

sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu

CASES = [(3, 2, (20, 4, 4), 128, 0), (3, 2, (20, 4, 4), 128, 24), (3, 2, (20, 4, 4), 16, 60), (3, 1, (24, 6, 6), 64, 24),
         (2, 2, (40, 8), 96, 12), (2, 1, (40, 8), 24, 10)]


@pytest.mark.parametrize("N,deg,sizes,aggregates,fine", CASES)
def test_preconditioner_operator_matches_numpy_restatement(lib_built, N, deg, sizes, aggregates, fine):
    import meshfem_b200
    import emulate_multilevel as em
    sim, fixed, vals, f = cantilever_problem(N, deg, sizes)
    op, info = em.full_operator(sim, fixed, aggregates, fine)
    rng = np.random.default_rng(5)
    n = sim.mesh.num_nodes * N
    with meshfem_b200.Handle(0, coarse_aggregates=aggregates, coarse_fine_nodes=fine) as h:
        h.set_mesh(N, deg, sim.mesh.nodes, sim.mesh.elem_nodes)
        h.set_material(sim.D)
        h.assemble()
        h.fix_variables(fixed, vals)
        r1, r2 = rng.standard_normal(n), rng.standard_normal(n)
        z1, rz1 = h.apply_preconditioner(r1)
        z2, rz2 = h.apply_preconditioner(r2)
        zb, _ = h.apply_preconditioner(np.asarray(f).reshape(-1))
    e1, erz1 = op(r1)
    e2, erz2 = op(r2)
    eb, _ = op(np.asarray(f).reshape(-1))
    scale = np.linalg.norm(e1)
    assert np.linalg.norm(z1.reshape(-1) - e1) < 1e-8 * scale, (info, np.linalg.norm(z1.reshape(-1) - e1) / scale)
    assert np.linalg.norm(z2.reshape(-1) - e2) < 1e-8 * np.linalg.norm(e2)
    assert np.linalg.norm(zb.reshape(-1) - eb) < 1e-8 * np.linalg.norm(eb)
    assert abs(rz1 - erz1) < 1e-8 * abs(erz1) and abs(rz2 - erz2) < 1e-8 * abs(erz2)
    free = np.ones(n, bool); free[fixed] = False
    assert abs(rz1 - (r1 * free) @ z1.reshape(-1)) < 1e-9 * abs(rz1)                     # r.z is what the PCG needs
    assert abs(r2 @ z1.reshape(-1) - r1 @ z2.reshape(-1)) < 1e-9 * abs(rz1)             # symmetric operator
