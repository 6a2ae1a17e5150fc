"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the
same inputs.  Tolerances: assembled values 1e-13 relative (Frobenius), operator applications
1e-12, displacements 1e-8 relative L2 (north-star gate: 1e-6)."""
import numpy as np
import pytest
import scipy.sparse as sp

import meshfem_oracle as orc
from util import ORTHO, cantilever_problem, grid_mesh, rel_l2

pytestmark = pytest.mark.gpu

CASES = [(2, 1, (6, 4)), (2, 2, (5, 3)), (3, 1, (4, 3, 2)), (3, 2, (3, 2, 2))]


def _material(N, kind, ne=None, seed=0):
    if kind == "iso":
        return orc.isotropic_D(N, 200.0, 0.35)
    if kind == "ortho":
        return orc.material_from_json(3, ORTHO) if N == 3 else orc.orthotropic_D2(200.0, 120.0, 0.18, 60.0)
    rng = np.random.default_rng(seed)
    F = orc.flat_len(N)
    A = rng.normal(size=(ne, F, F))
    return np.einsum("eij,ekj->eik", A, A) + 3 * np.eye(F)


@pytest.fixture(scope="module")
def mfem(lib_built):
    import meshfem_b200
    return meshfem_b200


def _handle(mfem, mesh, D, **opt):
    h = mfem.Handle(0, **opt)
    h.set_mesh(mesh.N, mesh.deg, mesh.nodes, mesh.elem_nodes)
    h.set_material(D)
    return h


@pytest.mark.parametrize("N,deg,sizes", CASES)
@pytest.mark.parametrize("mat", ["iso", "ortho", "perelem"])
@pytest.mark.parametrize("reorder", [0, 1])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_assembled_matrix_matches_oracle(mfem, N, deg, sizes, mat, reorder, mode):
    mesh = grid_mesh(N, deg, sizes)
    D = _material(N, mat, mesh.num_elements)
    with _handle(mfem, mesh, D, reorder=reorder, assembly=mode) as h:
        h.assemble()
        K = h.get_matrix().tocsr()
        vol = h.volumes()
    Kref = orc.stiffness_matrix(mesh, D)
    assert np.allclose(vol, mesh.vol, rtol=1e-14, atol=0)
    assert K.shape == Kref.shape
    diff = (K - Kref)
    assert sp.linalg.norm(diff) <= 1e-13 * sp.linalg.norm(Kref)
    # every significant oracle entry is structurally present in the block pattern
    big = abs(Kref) > 1e-9 * abs(Kref).max()
    assert (big.astype(np.int8) - big.multiply(K != 0).astype(np.int8)).nnz == 0


@pytest.mark.parametrize("N,deg,sizes", CASES)
def test_assembly_is_bit_reproducible(mfem, N, deg, sizes):
    mesh = grid_mesh(N, deg, sizes)
    D = _material(N, "ortho")
    vals = []
    for _ in range(2):
        with _handle(mfem, mesh, D) as h:
            h.assemble()
            vals.append(h.get_bsr()[2])
    assert np.array_equal(vals[0], vals[1])


@pytest.mark.parametrize("N,deg,sizes", CASES + [(3, 2, (9, 4, 3))])
@pytest.mark.parametrize("lanes,kernel", [(0, 0), (0, 1), (8, 1), (16, 1), (32, 1), (0, 2), (32, 3), (0, 4), (8, 4), (16, 4), (32, 4),
                                          (8, 11), (16, 11), (32, 11)])
def test_spmv_and_apply_K(mfem, N, deg, sizes, lanes, kernel):
    """kernel 0 = auto, 1 = direct-load SpMV (8/16/32 lanes per row), 2 = TMA-ring SpMV,
    3 = index-pipelined SpMV (32 lanes), 4 = symmetric SpMV (upper tails + atomic adds); 11 = kernel 1 with the L2
    prefetch of the next row (option spmv_prefetch)."""
    opts = dict(spmv_prefetch=1) if kernel == 11 else dict(spmv_prefetch=0)
    kernel = 1 if kernel == 11 else kernel
    mesh = grid_mesh(N, deg, sizes)
    D = _material(N, "ortho")
    rng = np.random.default_rng(5)
    x = rng.normal(size=(mesh.num_nodes, N))
    Kref = orc.stiffness_matrix(mesh, D)
    with _handle(mfem, mesh, D, spmv_lanes=lanes, spmv_kernel=kernel, **opts) as h:
        h.assemble()
        y = h.spmv(x)
        z = h.apply_K(x)
    yref = (Kref @ x.reshape(-1)).reshape(-1, N)
    assert rel_l2(y, yref) < 1e-13
    assert rel_l2(z, orc.apply_stiffness_matrix(mesh, D, x)) < 1e-13
    assert rel_l2(z, y) < 1e-12


@pytest.mark.parametrize("N,deg,sizes", CASES)
def test_periodic_dof_map_assembly(mfem, N, deg, sizes):
    """Nodes sharing a DoF (PeriodicCondition): K must be assembled in DoF space exactly as
    LinearElasticity.hh:1418-1430 does, including two local nodes of one element mapping to the
    same DoF."""
    mesh = grid_mesh(N, deg, sizes)
    rng = np.random.default_rng(3)
    nn = mesh.num_nodes
    dof = np.arange(nn)
    merged = rng.choice(nn, size=max(2, nn // 5), replace=False)
    dof[merged] = merged[0] if False else rng.choice(merged, size=merged.size)   # random identifications
    _, dof = np.unique(dof, return_inverse=True)
    nd = int(dof.max() + 1)
    D = _material(N, "iso")
    h = mfem.Handle(0)
    h.set_mesh(N, deg, mesh.nodes, mesh.elem_nodes, dof_for_node=dof, n_dofs=nd)
    h.set_material(D)
    h.assemble()
    K = h.get_matrix().tocsr()
    eps = rng.normal(size=orc.flat_len(N))
    load = h.const_strain_load(eps)
    h.close()
    Kref = orc.stiffness_matrix(mesh, D, dof, nd)
    assert sp.linalg.norm(K - Kref) <= 1e-13 * sp.linalg.norm(Kref)
    assert rel_l2(load, orc.constant_strain_load(mesh, D, eps, dof, nd)) < 1e-13


@pytest.mark.parametrize("N,deg,sizes", CASES)
@pytest.mark.parametrize("mat", ["iso", "perelem"])
def test_loads_and_strain_stress(mfem, N, deg, sizes, mat):
    mesh = grid_mesh(N, deg, sizes)
    D = _material(N, mat, mesh.num_elements)
    rng = np.random.default_rng(11)
    eps = rng.normal(size=orc.flat_len(N))
    u = rng.normal(size=(mesh.num_nodes, N))
    with _handle(mfem, mesh, D) as h:
        load = h.const_strain_load(eps)
        strain, stress = h.avg_strain_stress(u)
    assert rel_l2(load, orc.constant_strain_load(mesh, D, eps)) < 1e-13
    e_ref, s_ref = orc.average_strain_stress(mesh, D, u)
    assert rel_l2(strain, e_ref) < 1e-13
    assert rel_l2(stress, s_ref) < 1e-13


@pytest.mark.parametrize("N,deg,sizes", [(2, 1, (20, 4)), (2, 2, (10, 2)), (3, 1, (10, 2, 2)), (3, 2, (10, 2, 2))])
@pytest.mark.parametrize("reorder,kernel", [(0, 1), (1, 1), (1, 2), (1, 3), (1, 0), (1, 4), (0, 4)])
def test_cantilever_displacements_match_direct_solve(mfem, N, deg, sizes, reorder, kernel):
    sim, fixed, vals, f = cantilever_problem(N, deg, sizes)
    u_ref = sim.solve(f)
    with _handle(mfem, sim.mesh, sim.D, reorder=reorder, spmv_kernel=kernel, spmv_lanes=32 if kernel == 3 else 0) as h:
        h.assemble()
        h.fix_variables(fixed, vals)
        u, info = h.solve(f, rtol=1e-12, return_info=True)
    assert info[0]["converged"]
    err = rel_l2(u, u_ref)
    assert err < 1e-8, f"rel L2 {err}, iterations {info[0]['iterations']}"


@pytest.mark.parametrize("N,deg,sizes", [(3, 2, (6, 2, 2)), (3, 1, (8, 3, 2)), (2, 2, (8, 3)), (2, 1, (12, 4))])
def test_batched_pcg_matches_sequential_and_direct(mfem, N, deg, sizes):
    """flatLen(N) right-hand sides in one call run as ONE batched PCG (SpMM, csrc/solver_multi.inl): every
    system must reproduce the one-at-a-time solve and the oracle's direct solve, whatever its scale --
    including a zero load vector and systems that finish at different iterations (frozen while the
    others continue)."""
    sim, fixed, vals, f = cantilever_problem(N, deg, sizes)
    F = orc.flat_len(N)
    rng = np.random.default_rng(3)
    vals = 1e-3 * rng.normal(size=vals.shape)
    rhs = np.stack([f * 1.0] + [rng.normal(size=f.shape) * 10.0 ** (-2 * k) for k in range(1, F)])
    rhs[F - 1] = 0.0
    rhs[1, :, :] *= np.linspace(0, 1, f.shape[0])[:, None] ** 8        # smooth, fast-converging load
    K = sim.stiffness()
    ref = [orc.solve_fixed(K, r.reshape(-1), fixed, vals).reshape(-1, N) for r in rhs]
    out = {}
    for batch, kern in ((1, 0), (1, 2), (0, 0)):           # full-warp SpMM, half-warp split SpMM, one at a time
        with _handle(mfem, sim.mesh, sim.D, batch_rhs=batch, spmm_kernel=kern) as h:
            h.assemble()
            h.fix_variables(fixed, vals)
            out[batch, kern] = h.solve(rhs, rtol=1e-12, return_info=True)
    us, iseq = out[0, 0]
    assert all(i["converged"] for i in iseq)
    for kern in (0, 2):
        ub, ib = out[1, kern]
        assert all(i["converged"] for i in ib)
        for k in range(F):
            scale = max(np.linalg.norm(ref[k]), 1e-300)
            assert np.linalg.norm(ub[k] - ref[k]) / scale < 1e-8, k
            assert np.linalg.norm(ub[k] - us[k]) / scale < 1e-9, k
            assert abs(ib[k]["iterations"] - iseq[k]["iterations"]) <= 2, (k, ib[k], iseq[k])
        assert len({i["iterations"] for i in ib}) > 1      # the systems did finish at different iterations


@pytest.mark.parametrize("N,deg,sizes", [(3, 2, (5, 2, 2)), (2, 1, (9, 4))])
@pytest.mark.parametrize("upper", [True, False])
def test_external_matrix_spsd_system(mfem, N, deg, sizes, upper):
    """Seam S2 (SPSDSystem(K), SparseMatrices.hh:2321-2348): a matrix assembled elsewhere -- here the
    oracle's K as triplets, upper triangle with REPEATED entries as the reference produces them
    (LinearElasticity.hh:1408-1466) or full -- is summed, laid out on the device and solved with fixed
    variables; SpMV, export and the solution must match scipy."""
    sim, fixed, vals, f = cantilever_problem(N, deg, sizes)
    vals = 1e-3 * np.cos(np.arange(vals.size))
    K = sim.stiffness().tocsr()
    if upper:
        I, J, W = orc.assemble_upper_triplets(sim.mesh, sim.D)      # unsummed, i <= j
        Kin = (K.shape[0], I, J, W)
    else:
        Kin = K
    rng = np.random.default_rng(9)
    x = rng.normal(size=K.shape[0])
    with mfem.Handle(0) as h:
        h.set_matrix(Kin, block_dim=N, upper_triangle_only=upper)
        y = h.spmv(x.reshape(-1, N)).reshape(-1)
        assert rel_l2(y, K @ x) < 1e-13
        assert abs(h.get_matrix() - K).max() <= 1e-13 * abs(K).max()
        h.fix_variables(fixed, vals)
        u, info = h.solve(f, rtol=1e-12, return_info=True)
        assert info[0]["converged"]
        assert rel_l2(u, orc.solve_fixed(K, f.reshape(-1), fixed, vals)) < 1e-8
        with pytest.raises(mfem.MfemB200Error, match="set_matrix_triplets"):
            h.assemble()
    with mfem.Handle(0) as h:
        with pytest.raises(mfem.MfemB200Error, match="below the diagonal"):
            h.set_matrix((4, [1], [0], [1.0]), block_dim=2, upper_triangle_only=True)
        with pytest.raises(mfem.MfemB200Error, match="multiple of block_dim"):
            h.set_matrix((5, [0], [0], [1.0]), block_dim=2)


def test_nonzero_dirichlet_values_and_multiple_rhs(mfem):
    """fixVariables with non-zero values moves K_fc u_c to the RHS (SparseMatrices.hh:2457-2470)."""
    sim, fixed, vals, f = cantilever_problem(3, 2, (4, 2, 2), D=orc.material_from_json(3, ORTHO))
    rng = np.random.default_rng(2)
    vals = 1e-2 * rng.normal(size=vals.shape)
    K = sim.stiffness()
    f2 = rng.normal(size=f.shape)
    ref = [orc.solve_fixed(K, ff.reshape(-1), fixed, vals) for ff in (f, f2)]
    with _handle(mfem, sim.mesh, sim.D) as h:
        h.assemble()
        h.fix_variables(fixed[: fixed.size // 2], vals[: fixed.size // 2])      # cumulative calls
        h.fix_variables(fixed[fixed.size // 2:], vals[fixed.size // 2:])
        u = h.solve(np.concatenate([f.reshape(-1), f2.reshape(-1)]), rtol=1e-12)
    u = u.reshape(2, -1)
    for k in range(2):
        assert rel_l2(u[k], ref[k]) < 1e-8
        assert np.array_equal(u[k][fixed], vals)


def test_error_behaviour(mfem):
    mesh = grid_mesh(3, 1, (2, 2, 2))
    D = orc.isotropic_D(3, 1.0, 0.3)
    # negatively oriented element -> the reference's constructor error (LinearElasticity.hh:465-472)
    bad = mesh.elem_nodes.copy()
    bad[0, [0, 1]] = bad[0, [1, 0]]
    h = mfem.Handle(0)
    with pytest.raises(mfem.MfemB200Error) as ei:
        h.set_mesh(3, 1, mesh.nodes, bad)
    assert ei.value.status == -3 and "negatively oriented" in str(ei.value)
    h.close()
    with _handle(mfem, mesh, D) as h:
        with pytest.raises(mfem.MfemB200Error) as ei:   # "No system to solve"
            h.solve(np.zeros(mesh.num_nodes * 3))
        h.assemble()
        h.fix_variables([0, 1, 2])
        with pytest.raises(mfem.MfemB200Error) as ei:   # SparseMatrices.hh:2432
            h.fix_variables([2])
        assert ei.value.status == -4 and "Variable already fixed." in str(ei.value)
    with _handle(mfem, mesh, D) as h:                    # singular system (no Dirichlet): PCG must not "converge"
        h.assemble()
        rng = np.random.default_rng(0)
        with pytest.raises(mfem.MfemB200Error) as ei:
            h.solve(rng.normal(size=mesh.num_nodes * 3), rtol=1e-12, max_iters=300)
        assert ei.value.status in (-6, -7, -8)


def test_launch_counter_and_timers(mfem):
    sim, fixed, vals, f = cantilever_problem(3, 1, (6, 2, 2))
    with _handle(mfem, sim.mesh, sim.D) as h:
        h.assemble()
        h.fix_variables(fixed, vals)
        n0 = h.launch_count()
        h.solve(f, rtol=1e-10)
        assert h.launch_count() > n0 + 3
        assert h.timer("Assemble System") > 0 and h.timer("Elasticity Solve") > 0
        assert h.time_spmv(5) > 0
