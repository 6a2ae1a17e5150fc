"""GPU: the discrete shape-derivative element kernels (csrc/shape.cu: applyDeltaStiffnessMatrix, deltaConstantStrainLoad,
deltaAverageStrainField -- LinearElasticity.hh:1301-1375) through the C ABI against the oracle's formulas, which are
themselves pinned by central finite differences of K(p) u, load(p) and strain(p) in tests/test_shape_derivatives.py.
Plain and periodic DoF maps, constant and per-element materials, all four element types."""
import numpy as np
import pytest

import meshfem_oracle as orc
from util import ORTHO, grid_mesh

pytestmark = pytest.mark.gpu

CASES = [(2, 1, (6, 4)), (2, 2, (5, 3)), (3, 1, (4, 3, 2)), (3, 2, (4, 2, 2))]


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max()


@pytest.mark.parametrize("N,deg,sizes", CASES)
@pytest.mark.parametrize("material", ["constant", "per-element"])
@pytest.mark.parametrize("periodic", [False, True])
def test_shape_derivative_kernels_match_oracle(lib_built, N, deg, sizes, material, periodic):
    import meshfem_b200
    mesh = grid_mesh(N, deg, sizes)
    rng = np.random.default_rng(11)
    D = orc.isotropic_D(N, 200.0, 0.35) if N == 2 else orc.material_from_json(3, ORTHO)
    if material == "per-element":
        D = np.stack([D * s for s in rng.uniform(0.5, 2.0, size=mesh.num_elements)])
    dof, nd = (None, None)
    if periodic:
        dof, nd, _ = orc.periodic_condition(mesh)
    nv = mesh.vertices.shape[0]
    dp = rng.standard_normal((nv, N)) * 0.1
    u = rng.standard_normal((mesh.num_nodes, N))
    du = rng.standard_normal((mesh.num_nodes, N))
    F = N * (N + 1) // 2
    eps = rng.standard_normal(F)
    with meshfem_b200.Handle(0) as h:
        h.set_mesh(N, deg, mesh.nodes, mesh.elem_nodes, dof_for_node=dof, n_dofs=nd)
        h.set_material(D)
        dKu = h.apply_delta_K(u, dp)
        dl = h.delta_const_strain_load(eps, dp)
        ds = h.delta_avg_strain(u, du, dp)
        with pytest.raises(meshfem_b200.MfemB200Error, match="per-vertex"):
            h.apply_delta_K(u, dp[: nv // 2])
    assert _rel(dKu, orc.apply_delta_stiffness_matrix(mesh, D, u, dp, dof, nd)) < 1e-12
    assert _rel(dl, orc.delta_constant_strain_load(mesh, D, eps, dp, dof, nd)) < 1e-12
    assert _rel(ds, orc.delta_average_strain_field(mesh, u, du, dp)) < 1e-12
