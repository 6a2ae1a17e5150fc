"""Discrete shape derivatives (LinearElasticity.hh:1286-1373 and
PeriodicHomogenization.hh:383-563 of the reference).  Two layers of evidence:
 * the oracle's formulas are pinned by central finite differences of the oracle's own K(p) u, load(p), strain(p),
   w(p) and Ch(p) on perturbed meshes (theory KAT, O(h^2) truncation);
 * the host C++ (include/MeshFEM/ShapeDerivatives.hh, Simulator / PeriodicHomogenization functions) matches the oracle
   to rounding."""
import numpy as np
import pytest

import meshfem_oracle as orc

ORTHO = {"type": "orthotropic", "young": [200, 120, 80], "poisson": [0.3, 0.2, 0.12, 0.3, 0.3, 0.18], "shear": [45, 35, 60]}
CASES = [(2, 4, 2, 1), (2, 4, 2, 2), (3, 4, 2, 1), (3, 4, 2, 2)]


@pytest.fixture(scope="module")
def hostlib(lib_built):
    from meshfem_b200 import hostlib as hl
    return hl


def _setup(hostlib, N, n, hole, deg):
    raw = hostlib.perforated_cell(N, n, hole)
    V, T = raw.arrays()
    V = V[:, :N]
    D = orc.isotropic_D(N, 200.0, 0.35) if N == 2 else orc.material_from_json(3, ORTHO)
    sim = orc.Simulator(N, deg, V, T)
    sim.set_material(D)
    w = orc.solve_cell_problems(sim)
    rng = np.random.default_rng(5)
    inner = ((V > 1e-9) & (V < 1 - 1e-9)).all(axis=1)          # keep the cell faces (periodicity, |Y|) in place
    dp = rng.standard_normal(V.shape) * inner[:, None] * 0.1
    u = rng.standard_normal((sim.mesh.num_nodes, N))
    return raw, V, T, D, sim, w, dp, u


def _perturbed(N, deg, V, T, D, dp, eps):
    s2 = orc.Simulator(N, deg, V, T)
    s2.set_material(D)
    dof, nd, pbe = orc.periodic_condition(s2.mesh)
    s2.set_periodic(dof, nd, pbe)
    s2.apply_no_rigid_motion_constraint(); s2.set_use_pin_no_rigid_translation_constraint(True)
    orc.set_node_positions(s2.mesh, V + eps * dp)
    return s2


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max()


@pytest.mark.parametrize("N,n,hole,deg", CASES)
def test_oracle_shape_derivatives_match_finite_differences(hostlib, N, n, hole, deg):
    raw, V, T, D, sim, w, dp, u = _setup(hostlib, N, n, hole, deg)
    m = sim.mesh
    F = orc.flat_len(N)
    h = 1e-5
    sp, sm = _perturbed(N, deg, V, T, D, dp, h), _perturbed(N, deg, V, T, D, dp, -h)

    def Ku(s_):
        f = orc.apply_stiffness_matrix(s_.mesh, D, u)
        out = np.zeros((s_.num_dofs(), N)); np.add.at(out, s_.dof_for_node, f)
        return out
    assert _rel((Ku(sp) - Ku(sm)) / (2 * h), orc.apply_delta_stiffness_matrix(m, D, u, dp, sim.dof_for_node, sim.num_dofs())) < 1e-7
    e = np.array([0.3, -0.2, 0.5, 0.1, 0.7, -0.4])[:F]
    assert _rel((sp.constant_strain_load(e) - sm.constant_strain_load(e)) / (2 * h),
                orc.delta_constant_strain_load(m, D, e, dp, sim.dof_for_node, sim.num_dofs())) < 1e-7

    def cell(s_):
        ww = [s_.solve(s_.constant_strain_load(-orc.canonical_basis(N, i))) for i in range(F)]
        return orc.homogenized_tensor_displacement_form(s_, ww, 1.0), ww
    Ep, wp = cell(sp); Em, wm = cell(sm)
    dCh = orc.homogenized_tensor_discrete_differential(sim, w)
    an = np.einsum("vcfg,vc->fg", dCh, dp)
    assert _rel((Ep - Em) / (2 * h), an) < 1e-7
    assert np.abs(an - an.T).max() < 1e-12 * np.abs(an).max()
    dw = orc.delta_fluctuation_displacements(sim, w, dp)
    assert _rel((np.array(wp) - np.array(wm)) / (2 * h), np.array(dw)) < 1e-7
    fd = (orc.average_strain_stress(sp.mesh, D, wp[1])[0] - orc.average_strain_stress(sm.mesh, D, wm[1])[0]) / (2 * h)
    assert _rel(fd, orc.delta_average_strain_field(m, w[1], dw[1], dp)) < 1e-7
    # rigid translation of ALL vertices changes nothing
    t = np.tile(np.array([0.3, -0.7, 0.2])[:N], (V.shape[0], 1))
    assert np.abs(orc.apply_delta_stiffness_matrix(m, D, u, t)).max() < 1e-10 * np.abs(Ku(sim)).max()
    assert np.abs(np.einsum("vcfg,vc->fg", dCh, t)).max() < 1e-10 * np.abs(Ep).max()


@pytest.mark.parametrize("N,n,hole,deg", CASES)
@pytest.mark.parametrize("periodic", [False, True])
def test_host_shape_derivatives_match_oracle(hostlib, N, n, hole, deg, periodic):
    raw, V, T, D, sim, w, dp, u = _setup(hostlib, N, n, hole, deg)
    m = sim.mesh
    F = orc.flat_len(N)
    rng = np.random.default_rng(11)
    du = rng.standard_normal(u.shape)
    e = np.array([0.3, -0.2, 0.5, 0.1, 0.7, -0.4])[:F]
    dofs = (sim.dof_for_node, sim.num_dofs()) if periodic else (None, None)
    r = raw.shape_derivatives(deg, D, u, du, e, dp, w_ij=np.array(w), periodic=periodic, num_dofs=dofs[1])
    assert _rel(r["dKu"], orc.apply_delta_stiffness_matrix(m, D, u, dp, *dofs)) < 1e-12
    assert _rel(r["dload"], orc.delta_constant_strain_load(m, D, e, dp, *dofs)) < 1e-12
    assert _rel(r["dstrain"], orc.delta_average_strain_field(m, u, du, dp)) < 1e-12
    assert _rel(r["dCh"], orc.homogenized_tensor_discrete_differential(sim, w)) < 1e-11
