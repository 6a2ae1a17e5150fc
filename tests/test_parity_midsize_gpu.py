"""GPU: parity AT SCALE -- the largest problem the oracle's direct solver handles (quadratic-tet cantilever 40x8x8:
61,440 elements, 92,785 nodes, 278,355 DoF; golden field tests/golden/cantilever_40x8x8_deg2.npz from
tests/golden/make_midsize_golden.py, a 3-minute sparse LU) against the device path through the C ABI with every
preconditioner configuration the bench uses: block-Jacobi, the multilevel method with explicit sizes, and the library's
automatic choice.  Bar: full-vector relative L2 <= 1e-8 (the north star's gate is 1e-6); also the true residual through
the independent matrix-free element-wise K u (mfem_b200_apply_K), and agreement between rtol 1e-8 (the bench's) and the
converged answer."""
import os
import sys

import numpy as np
import pytest

from util import GOLDEN, ROOT, rel_l2

sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu

FIXTURE = os.path.join(GOLDEN, "cantilever_40x8x8_deg2.npz")


@pytest.fixture(scope="module")
def problem(lib_built):
    import workloads as wl
    gold = np.load(FIXTURE)
    sizes = tuple(int(x) for x in gold["sizes"])
    m = wl.grid_femmesh(sizes, 2)
    fixed, vals, f = wl.cantilever_inputs(m)
    return m, wl.material("iso"), fixed, vals, f, gold["u"]


@pytest.mark.parametrize("coarse,fine", [(0, 0), (256, 0), (256, 32), (-1, 64)])
def test_midsize_cantilever_matches_direct_solve(problem, coarse, fine):
    import meshfem_b200
    m, D, fixed, vals, f, u_ref = problem
    with meshfem_b200.Handle(0, coarse_aggregates=coarse, coarse_fine_nodes=fine) as h:
        h.set_mesh(3, 2, m.nodes, m.elem_nodes)
        h.set_material(D)
        h.assemble()
        h.fix_variables(fixed, vals)
        u, info = h.solve(f, rtol=1e-11, return_info=True)
        u8, info8 = h.solve(f, rtol=1e-8, return_info=True)
        Ku = h.apply_K(u.reshape(-1, 3))
    assert info[0]["converged"] and info8[0]["converged"]
    err = rel_l2(u, u_ref)
    assert err < 1e-8, (coarse, fine, err, info[0]["iterations"])
    assert rel_l2(u8, u_ref) < 1e-6, rel_l2(u8, u_ref)                       # the bench's tolerance meets the north-star gate
    free = np.ones(u.size, bool); free[np.asarray(fixed)] = False
    fr = np.asarray(f).reshape(-1)
    assert np.linalg.norm((fr - Ku.reshape(-1))[free]) <= 1e-9 * np.linalg.norm(fr[free])
    if coarse:      # the aggregation levels must pay for themselves: block-Jacobi needs ~1100 iterations here
        assert info8[0]["iterations"] < 450, info8[0]["iterations"]
