"""Lagrange-multiplier rows of the elasticity system (assembleConstrainedSystem
LinearElasticity.hh:1201-1249 with m_appendInfinitesimalRotationMatrix :1530-1568 / m_appendTranslationMatrix
:1571-1593).  The reference solves the saddle-point matrix with UMFPACK (oracle: solve_constrained, sparse LU of
the same matrix); the product resolves the rows on the host around the SPSD device solve
(include/MeshFEM/RigidMotionConstraints.hh).  These CPU tests run that C++ algebra with the SPSD solve supplied
by the test (dense least squares + a deliberate null-space component) and compare with the oracle."""
import json

import numpy as np
import pytest

import meshfem_oracle as orc

FACE_MAX_X = {"minCorner": [0.999, -0.01, -0.01], "maxCorner": [1.001, 1.01, 1.01]}
FACE_MIN_X = {"minCorner": [-0.001, -0.01, -0.01], "maxCorner": [0.001, 1.01, 1.01]}
FACE_MIN_Y = {"minCorner": [-0.01, -0.001, -0.01], "maxCorner": [1.01, 0.001, 1.01]}


@pytest.fixture(scope="module")
def hostlib(lib_built):
    from meshfem_b200 import hostlib as hl
    return hl


def _oracle_sim(N, deg, sizes, bc, periodic=False, pin=False):
    V, T = orc.grid_simplices(list(sizes))
    sim = orc.Simulator(N, deg, V, T)
    sim.set_material(orc.isotropic_D(N, 200.0, 0.35))
    if periodic:
        dof, nd, pbe = orc.periodic_condition(sim.mesh)
        sim.set_periodic(dof, nd, pbe)
        sim.apply_no_rigid_motion_constraint()
    sim.set_use_pin_no_rigid_translation_constraint(pin)
    if bc:
        conds, nrm, pps, pins = orc.read_boundary_conditions(N, bc, sim.mesh.bbox_min, sim.mesh.bbox_max)
        sim.apply_translation_pins(pins)
        sim.apply_boundary_conditions(conds)
        if nrm:
            sim.apply_no_rigid_motion_constraint()
    return sim


def _spsd_solver(K, fixed, vals, null_mode):
    """Any solution of the consistent, possibly singular system K_ff u_f = b_f - K_fc u_c: the least-norm one plus
    a multiple of a null-space mode (what an iterative solver is free to return)."""
    n = K.shape[0]
    Kd = K.toarray()
    free = np.ones(n, bool); free[fixed] = False

    def solve(B):
        out = np.zeros_like(B)
        for k in range(B.shape[0]):
            u = np.zeros(n); u[fixed] = vals
            b = B[k][free] - Kd[free][:, ~free] @ u[~free]
            u[free] = np.linalg.lstsq(Kd[free][:, free], b, rcond=1e-12)[0]
            if null_mode is not None:
                u[free] += 0.37 * null_mode[free]
            out[k] = u
        return out
    return solve


def _check(hostlib, N, deg, sizes, bc, periodic=False, pin=False, nrows=None, extra_rhs=0):
    sim = _oracle_sim(N, deg, sizes, bc, periodic, pin)
    fixed, vals, C, d = sim.constraints()
    r = hostlib.grid(list(sizes)).apply_bc(deg, json.dumps(bc) if bc else "", periodic=periodic, pin=pin)
    # the rows and the fixed variables are the reference's
    assert r["constraint_rows"].shape == C.shape and (nrows is None or C.shape[0] == nrows)
    assert np.array_equal(r["constraint_rows"], C) and np.array_equal(r["constraint_rhs"], d)
    assert np.array_equal(r["fixed_vars"], fixed) and np.array_equal(r["fixed_vals"], vals)
    K = sim.stiffness().tocsr()
    n = K.shape[0]
    rng = np.random.default_rng(7)
    F = [sim.neumann_load().reshape(-1)] + [rng.standard_normal(n) for _ in range(extra_rhs)]
    if not np.abs(F[0]).max() > 0:
        F[0] = rng.standard_normal(n)
    # the null-space component the stand-in solver adds must itself vanish on the fixed variables
    null_mode = None
    for mode in r["rigid_modes"]:
        if np.abs(mode[fixed]).max(initial=0.0) == 0.0:
            null_mode = mode
            break
    U, lam = hostlib.constrained_solve(r["constraint_rows"], r["constraint_rhs"], r["fixed_vars"], r["rigid_modes"],
                                       np.array(F), _spsd_solver(K, fixed, vals, null_mode))
    for k, f in enumerate(F):
        u_ref, lam_ref = orc.solve_constrained(K, C, d, f, fixed, vals)
        scale = np.abs(u_ref).max()
        assert np.abs(U[k] - u_ref).max() <= 1e-9 * scale, (k, np.abs(U[k] - u_ref).max() / scale)
        assert np.allclose(lam[k], lam_ref, rtol=1e-8, atol=1e-10 * max(1.0, np.abs(f).max()))
        assert np.abs(C @ U[k] - d).max() <= 1e-10 * max(1.0, scale * np.abs(C).max())
    return sim, U, lam


@pytest.mark.parametrize("N,deg,sizes", [(3, 1, (4, 2, 2)), (3, 2, (3, 2, 2)), (2, 1, (6, 3)), (2, 2, (5, 3))])
def test_no_rigid_motion_rows_unbalanced_load(hostlib, N, deg, sizes):
    """no_rigid_motion with a load that is NOT self-equilibrated: the multipliers carry the resultant."""
    val = [3.0, 1.0, 0.0][:N] + [0.0] * (3 - N)
    box = {k: v[:N] + [0.0] * 0 for k, v in FACE_MAX_X.items()} if N == 3 else {"minCorner": [0.999, -0.01], "maxCorner": [1.001, 1.01]}
    bc = {"no_rigid_motion": True, "regions": [{"type": "force", "value": val, "box%": box}]}
    sim, U, lam = _check(hostlib, N, deg, sizes, bc, nrows=6 if N == 3 else 3, extra_rhs=1)
    assert np.abs(lam[0]).max() > 1e-4                    # unbalanced: non-zero multipliers


def test_no_rigid_motion_balanced_load_is_uniaxial_patch(hostlib):
    """Opposite tractions on the two x faces: multipliers vanish and the solution is the exact uniaxial-stress
    state (linear displacement, reproduced exactly by the elements)."""
    t = 2.5
    bc = {"no_rigid_motion": True, "regions": [{"type": "traction", "value": [t, 0, 0], "box%": FACE_MAX_X},
                                               {"type": "traction", "value": [-t, 0, 0], "box%": FACE_MIN_X}]}
    sim, U, lam = _check(hostlib, 3, 2, (3, 2, 2), bc, nrows=6)
    assert np.abs(lam[0]).max() < 1e-10
    u = sim.dof_to_node_field(U[0])
    strain, stress = sim.average_strain_stress(u)
    assert np.allclose(stress[:, 0], t, rtol=1e-9) and np.abs(stress[:, 1:]).max() < 1e-9
    assert np.allclose(strain[:, 0], t / 200.0, rtol=1e-9) and np.allclose(strain[:, 1], -0.35 * t / 200.0, rtol=1e-9)


def test_no_rigid_motion_with_pinned_node(hostlib):
    """setUsePinNoRigidTranslationConstraint(true) without periodicity: three rotation rows + a pinned node
    (:1216-1218); the free rigid modes are the rotations about that node."""
    bc = {"no_rigid_motion": True, "regions": [{"type": "traction", "value": [0, 1.0, 0.5], "box%": FACE_MAX_X},
                                               {"type": "traction", "value": [0, -1.0, -0.5], "box%": FACE_MIN_X}]}
    _check(hostlib, 3, 1, (3, 3, 2), bc, pin=True, nrows=3, extra_rhs=1)


def test_unconstrained_translation_rows(hostlib):
    """Dirichlet conditions that leave translation components free and no pin: one translation row per free
    component (:1236-1238).  2D: y fixed on the bottom edge (removes the y translation and the rotation), x free."""
    bc = {"regions": [{"type": "dirichlety", "value": [0, "0.01*x", 0], "box%": {"minCorner": [-0.01, -0.001], "maxCorner": [1.01, 0.001]}},
                      {"type": "force", "value": [1.0, -2.0, 0], "box%": {"minCorner": [-0.01, 0.999], "maxCorner": [1.01, 1.001]}}]}
    sim, U, lam = _check(hostlib, 2, 2, (5, 3), bc, nrows=1, extra_rhs=1)
    assert abs(lam[0][0]) > 1e-4                          # the x resultant of the load goes to the multiplier


@pytest.mark.parametrize("sizes,deg", [((3, 3, 3), 1), ((4, 4), 2)])
def test_periodic_translation_rows(hostlib, sizes, deg):
    """Periodic conditions + no-rigid-motion without the pin: N translation rows on the periodic DoFs, no rotation
    rows (:1539-1540)."""
    N = len(sizes)
    sim, U, lam = _check(hostlib, N, deg, sizes, None, periodic=True, pin=False, nrows=N, extra_rhs=2)


def test_rows_that_do_not_match_the_null_space_are_rejected(hostlib):
    """y fixed on one face only in 3D leaves the x,z translations AND the rotation about y free: two rows for three
    free modes -- the reference's matrix is singular there; here the mismatch is reported."""
    bc = {"regions": [{"type": "dirichlety", "value": [0, 0, 0], "box%": FACE_MIN_Y},
                      {"type": "force", "value": [1.0, 0, 0], "box%": FACE_MAX_X}]}
    sim = _oracle_sim(3, 1, (2, 2, 2), bc)
    fixed, vals, C, d = sim.constraints()
    r = hostlib.grid([2, 2, 2]).apply_bc(1, json.dumps(bc))
    assert C.shape[0] == 2
    K = sim.stiffness().tocsr()
    with pytest.raises(RuntimeError, match="rigid mode"):
        hostlib.constrained_solve(r["constraint_rows"], r["constraint_rhs"], r["fixed_vars"], r["rigid_modes"],
                                  sim.neumann_load().reshape(1, -1), _spsd_solver(K, fixed, vals, None))


def test_rigid_motion_constraint_with_rhs_oracle():
    """applyRigidMotionConstraint(u0) (:1069-1076): the solution has the rigid motion of u0."""
    sim = _oracle_sim(2, 1, (4, 3), {"no_rigid_motion": True, "regions": [
        {"type": "force", "value": [1.0, 0.5, 0], "box%": {"minCorner": [0.999, -0.01], "maxCorner": [1.001, 1.01]}}]})
    rng = np.random.default_rng(3)
    u0 = rng.standard_normal((sim.num_dofs(), 2))
    sim.apply_rigid_motion_constraint(u0)
    u = sim.solve()
    assert np.allclose(sim.rigid_inner_product(u), sim.rigid_inner_product(u0), rtol=1e-10, atol=1e-10)
    with pytest.raises(RuntimeError, match="Invalid rigid motion RHS"):
        sim.rigid_motion_rhs = np.zeros(2)
        sim.constraints()


def test_rows_on_top_of_a_definite_system_schur_complement(hostlib):
    """no_rigid_motion together with a clamped face: the Dirichlet conditions already remove every rigid mode (k = 0), the
    three rotation rows + three translation rows come on top -- the reference factorises that saddle point as well
    (SparseMatrices.hh:2332-2348); here the six rows are resolved by a Schur complement, six extra SPSD solves."""
    bc = {"no_rigid_motion": True, "regions": [{"type": "dirichlet", "value": [0, 0, 0], "box%": FACE_MIN_X},
                                               {"type": "force", "value": [0.5, -2.0, 1.0], "box%": FACE_MAX_X}]}
    sim, U, lam = _check(hostlib, 3, 1, (3, 2, 2), bc, nrows=6, extra_rhs=1)
    assert np.abs(lam[0]).max() > 1e-4
    # and the constrained solution is NOT the plain clamped-cantilever solution: the rows act
    fixed, vals, C, d = sim.constraints()
    u_plain = orc.solve_fixed(sim.stiffness(), sim.neumann_load().reshape(-1), fixed, vals)
    assert np.abs(U[0] - u_plain).max() > 1e-3 * np.abs(u_plain).max()


def test_more_rows_than_free_rigid_modes_mixed_case(hostlib):
    """2D, y fixed on the bottom edge (removes the y translation and the rotation, leaves the x translation: k = 1) and
    no_rigid_motion on top (one rotation row + two translation rows: m = 3): one multiplier direction is dictated by the
    null space, the other two by the Schur complement."""
    bc = {"no_rigid_motion": True,
          "regions": [{"type": "dirichlety", "value": [0, 0, 0], "box%": {"minCorner": [-0.01, -0.001], "maxCorner": [1.01, 0.001]}},
                      {"type": "force", "value": [1.0, -2.0, 0], "box%": {"minCorner": [-0.01, 0.999], "maxCorner": [1.01, 1.001]}}]}
    _check(hostlib, 2, 2, (5, 3), bc, nrows=3, extra_rhs=1)


def test_too_few_rows_is_still_rejected(hostlib):
    """Fewer rows than free rigid modes: singular for the reference too."""
    sim = _oracle_sim(3, 1, (2, 2, 2), {"no_rigid_motion": True, "regions": []})
    fixed, vals, C, d = sim.constraints()
    r = hostlib.grid([2, 2, 2]).apply_bc(1, json.dumps({"no_rigid_motion": True, "regions": []}))
    K = sim.stiffness().tocsr()
    with pytest.raises(RuntimeError, match="rigid mode"):
        hostlib.constrained_solve(r["constraint_rows"][:3], r["constraint_rhs"][:3], r["fixed_vars"], r["rigid_modes"],
                                  np.ones((1, K.shape[0])), _spsd_solver(K, fixed, vals, None))
