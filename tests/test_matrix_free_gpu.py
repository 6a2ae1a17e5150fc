"""GPU: the PCG's mesh-based (matrix-free) operator, csrc/matfree.inl -- y = K x evaluated element by element on the
degree-2 rule (what Simulator::applyStiffnessMatrix sums, LinearElasticity.hh:801-823) -- against the oracle's assembled
matrix and against the stored-matrix SpMV, and solves that iterate with it against the oracle's direct solve.
Tolerances: operator 1e-13 relative L2, displacements 1e-8 (north-star gate 1e-6)."""
import numpy as np
import pytest

import meshfem_oracle as orc
from util import ORTHO, cantilever_problem, grid_mesh, rel_l2

pytestmark = pytest.mark.gpu

CASES = [(2, 1, (6, 4)), (2, 2, (5, 3)), (3, 1, (4, 3, 2)), (3, 2, (3, 2, 2)), (3, 2, (9, 4, 3))]


@pytest.fixture(scope="module")
def mfem(lib_built):
    import meshfem_b200
    return meshfem_b200


def _material(N, kind, ne=None, seed=0):
    if kind == "iso":
        return orc.isotropic_D(N, 200.0, 0.35)
    if kind == "ortho":
        return orc.material_from_json(3, ORTHO) if N == 3 else orc.orthotropic_D2(200.0, 120.0, 0.18, 60.0)
    rng = np.random.default_rng(seed)
    F = orc.flat_len(N)
    A = rng.normal(size=(ne, F, F))
    return np.einsum("eij,ekj->eik", A, A) + 3 * np.eye(F)


@pytest.mark.parametrize("N,deg,sizes", CASES)
@pytest.mark.parametrize("mat", ["iso", "ortho", "perelem"])
@pytest.mark.parametrize("reorder", [0, 1])
def test_operator_matches_oracle_matrix(mfem, N, deg, sizes, mat, reorder):
    mesh = grid_mesh(N, deg, sizes)
    D = _material(N, mat, mesh.num_elements)
    rng = np.random.default_rng(17)
    x = rng.normal(size=(mesh.num_nodes, N))
    yref = (orc.stiffness_matrix(mesh, D) @ x.reshape(-1)).reshape(-1, N)
    with mfem.Handle(0, reorder=reorder, spmv_kernel=6) as h:
        h.set_mesh(N, deg, mesh.nodes, mesh.elem_nodes)
        h.set_material(D)
        h.assemble()
        y = h.spmv(x)
        y2 = h.spmv(x)
        ych = []
        for ch in (32, 64):                           # chunk sizes other than the default 128
            h.set_option("mf_chunk_elems", ch)
            ych.append(h.spmv(x))
        h.set_option("mf_chunked", 0)                 # one slot per (element, node) instead of per-chunk partial sums
        y3 = h.spmv(x)
        h.set_option("mf_elem_order", 1)              # ... with the elements in DoF order instead of the caller's
        y4 = h.spmv(x)
        h.set_option("spmv_kernel", 0)
        ys = h.spmv(x)
    assert rel_l2(y, yref) < 1e-13
    assert rel_l2(y3, yref) < 1e-13
    assert rel_l2(y4, yref) < 1e-13
    assert rel_l2(ych[0], yref) < 1e-13 and rel_l2(ych[1], yref) < 1e-13
    assert rel_l2(y, ys) < 1e-13
    assert np.array_equal(y, y2)                      # no atomics: bit-reproducible


@pytest.mark.parametrize("chunked,pad,lanes,order,policy", [(1, 0, 0, 0, 3), (1, 1, 8, 1, 1), (1, 0, 4, 0, 0), (1, 0, 1, 0, 3), (0, 0, 1, 0, 3), (0, 1, 4, 1, 1), (0, 0, 4, 1, 1),
                                                            (0, 1, 8, 1, 1), (0, 0, 8, 1, 3), (0, 1, 4, 0, 1), (0, 0, 8, 0, 0), (0, 1, 8, 1, 2)])
def test_operator_layout_variants_in_a_solve(mfem, chunked, pad, lanes, order, policy):
    """A/B variants of the operator (per-chunk partial sums or one slot per (element, node); packed 24-byte or padded
    32-byte slots; 4 or 8 lanes per DoF row in the in-loop gather; elements in DoF order or in the caller's order; L2
    policy of the slot loads): same solution."""
    sim, fixed, vals, f = cantilever_problem(3, 2, (9, 3, 2), D=orc.material_from_json(3, ORTHO))
    u_ref = sim.solve(f)
    with mfem.Handle(0, matrix_free=1, mf_chunked=chunked, mf_slot_pad=pad, mf_gather_lanes=lanes, mf_elem_order=order,
                     mf_gather_policy=policy, mf_chunk_warps=(12, 16, 20)[(lanes + policy) % 3],
                     mf_chunk_elems=(32, 64, 128)[(pad + lanes + order) % 3]) as h:
        h.set_mesh(3, 2, sim.mesh.nodes, sim.mesh.elem_nodes)
        h.set_material(sim.D)
        h.assemble()
        h.fix_variables(fixed, vals)
        u, info = h.solve(f, rtol=1e-12, return_info=True)
    assert info[0]["converged"] and rel_l2(u, u_ref) < 1e-8


@pytest.mark.parametrize("N,deg,sizes", [(2, 2, (5, 3)), (3, 2, (3, 2, 2))])
def test_operator_in_periodic_dof_space(mfem, N, deg, sizes):
    """Nodes identified by a DoF map (PeriodicCondition; LinearElasticity.hh:1418-1430), two local nodes of an element on
    the same DoF included: the element results are summed per DoF."""
    mesh = grid_mesh(N, deg, sizes)
    rng = np.random.default_rng(3)
    nn = mesh.num_nodes
    dof = np.arange(nn)
    merged = rng.choice(nn, size=max(2, nn // 5), replace=False)
    dof[merged] = rng.choice(merged, size=merged.size)
    _, dof = np.unique(dof, return_inverse=True)
    nd = int(dof.max() + 1)
    D = _material(N, "ortho")
    x = rng.normal(size=(nd, N))
    with mfem.Handle(0, spmv_kernel=6) as h:
        h.set_mesh(N, deg, mesh.nodes, mesh.elem_nodes, dof_for_node=dof, n_dofs=nd)
        h.set_material(D)
        h.assemble()
        y = h.spmv(x)
    yref = (orc.stiffness_matrix(mesh, D, dof, nd) @ x.reshape(-1)).reshape(-1, N)
    assert rel_l2(y, yref) < 1e-13


@pytest.mark.parametrize("N,deg,sizes", [(2, 1, (20, 4)), (2, 2, (10, 2)), (3, 1, (10, 2, 2)), (3, 2, (10, 2, 2))])
@pytest.mark.parametrize("coarse", [0, -1])
def test_solve_with_matrix_free_operator(mfem, N, deg, sizes, coarse):
    """PCG iterating with the mesh-based operator (forced on for every element type) reaches the oracle's direct solve, and
    the same solution as the PCG on the stored matrix, non-zero Dirichlet values included (b = f - K u_fix goes through
    the operator as well)."""
    sim, fixed, vals, f = cantilever_problem(N, deg, sizes)
    vals = np.array(vals, dtype=float)
    vals[::3] += 0.01
    out = {}
    for mf in (1, 0):
        with mfem.Handle(0, matrix_free=mf, coarse_aggregates=coarse) as h:
            h.set_mesh(N, deg, sim.mesh.nodes, sim.mesh.elem_nodes)
            h.set_material(sim.D)
            h.assemble()
            h.fix_variables(fixed, vals)
            u, info = h.solve(f, rtol=1e-12, return_info=True)
            tsec, parts, used = h.time_operator(2)
        assert info[0]["converged"]
        assert used == bool(mf)
        assert tsec > 0 and (not mf or (parts[0] > 0 and parts[1] > 0))
        out[mf] = (u, info[0]["iterations"])
    assert rel_l2(out[1][0], out[0][0]) < 1e-9
    assert abs(out[1][1] - out[0][1]) <= 2 + out[0][1] // 20          # same Krylov sequence up to rounding
    u_ref = orc.solve_fixed(sim.stiffness(), np.asarray(f, float).reshape(-1), fixed, vals)
    assert rel_l2(out[1][0], u_ref) < 1e-8


def test_auto_selects_operator_for_3d_quadratic_only(mfem):
    for N, deg, sizes, want in [(3, 2, (4, 2, 2), True), (3, 1, (4, 2, 2), False), (2, 2, (4, 4), False)]:
        sim, fixed, vals, f = cantilever_problem(N, deg, sizes)
        with mfem.Handle(0) as h:
            h.set_mesh(N, deg, sim.mesh.nodes, sim.mesh.elem_nodes)
            h.set_material(sim.D)
            h.assemble()
            h.fix_variables(fixed, vals)
            u = h.solve(f, rtol=1e-12)
            assert h.time_operator(1)[2] == want
        assert rel_l2(u, sim.solve(f)) < 1e-8
