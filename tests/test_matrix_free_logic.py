"""CPU: the index logic of the chunked matrix-free operator (csrc/matfree.inl k_mf_chunk / k_mf_gather, tables of
csrc/setup.cu build_mf_chunks) restated in numpy by tools/emulate_matrix_free.py -- CTA-wide stable sort of (DoF, slot)
pairs in blocked order, distinct DoFs / local indices / ranks / extents, per-chunk partial sums in sorted order, per-row
lists of partials -- against the oracle's assembled matrix (the reference quantity: Simulator::applyStiffnessMatrix,
LinearElasticity.hh:801-823).  The GPU tests (tests/test_matrix_free_gpu.py) check the kernels themselves."""
import os
import sys

import numpy as np
import pytest

import meshfem_oracle as orc
from util import ORTHO, ROOT, grid_mesh, rel_l2

sys.path.insert(0, os.path.join(ROOT, "tools"))
import emulate_matrix_free as emf  # noqa: E402


@pytest.mark.parametrize("N,deg,sizes", [(2, 2, (5, 3)), (3, 1, (4, 3, 2)), (3, 2, (5, 2, 2))])
@pytest.mark.parametrize("chunk", [32, 64, 128])
def test_chunked_operator_equals_assembled_matrix(N, deg, sizes, chunk):
    mesh = grid_mesh(N, deg, sizes)
    D = orc.material_from_json(3, ORTHO) if N == 3 else orc.orthotropic_D2(200.0, 120.0, 0.18, 60.0)
    Ke = emf.element_matrices(mesh, D)
    rng = np.random.default_rng(7)
    x = rng.normal(size=(mesh.num_nodes, N))
    tab = emf.build_chunk_tables(mesh.elem_nodes, mesh.num_nodes, chunk)
    mask = rng.random(mesh.num_nodes * N) < 0.1
    y, dot = emf.apply_operator(tab, Ke, x, N, fixed_mask=mask)
    yref = (orc.stiffness_matrix(mesh, D) @ x.reshape(-1))
    yref[mask] = 0.0
    assert rel_l2(y, yref) < 1e-13
    assert abs(dot - float(x.reshape(-1) @ yref)) <= 1e-12 * abs(dot)
    # table invariants the device kernel relies on
    ne, npe = mesh.elem_nodes.shape
    S = chunk * npe
    counts = np.diff(tab["chunk_base"])
    for b in range(tab["n_chunks"]):
        nu = counts[b]
        dofs = tab["chunk_dof"][b]
        assert np.all(np.diff(dofs[:nu]) > 0)                  # distinct, ascending
        assert np.all(dofs[nu:] == 0)                          # read before the count is known: must be valid addresses
        ptr = tab["csr_ptr"][b]
        n_valid = min(chunk, ne - b * chunk) * npe
        assert ptr[0] == 0 and ptr[nu] == n_valid and np.all(np.diff(ptr[:nu + 1]) > 0)
        # ranks of the valid slots are a permutation of 0 .. n_valid-1, grouped by local index
        slots = [i * chunk + t for t in range(min(chunk, ne - b * chunk)) for i in range(npe)]
        rk = tab["rank"][b, slots]
        assert sorted(rk) == list(range(n_valid))
        li = tab["local_idx"][b, slots]
        assert np.all((ptr[li] <= rk) & (rk < ptr[li + 1]))
        assert np.array_equal(dofs[li], mesh.elem_nodes[b * chunk: b * chunk + chunk].reshape(-1))
    assert tab["n_partials"] == counts.sum() < ne * npe        # the reduction the design is about
    assert tab["inc_ptr2"][0] == 0 and tab["inc_ptr2"][-1] == tab["n_partials"]


def test_chunked_operator_in_periodic_dof_space():
    """DoF map with identified nodes (PeriodicCondition): two local nodes of one element on the same DoF land in the same
    extent of the chunk and are both summed."""
    N, deg = 3, 2
    mesh = grid_mesh(N, deg, (3, 2, 2))
    rng = np.random.default_rng(3)
    nn = mesh.num_nodes
    dof = np.arange(nn)
    merged = rng.choice(nn, size=nn // 4, replace=False)
    dof[merged] = rng.choice(merged, size=merged.size)
    _, dof = np.unique(dof, return_inverse=True)
    nd = int(dof.max() + 1)
    D = orc.isotropic_D(3, 200.0, 0.35)
    elem_dof = dof[mesh.elem_nodes]
    assert any(len(set(r)) < len(r) for r in elem_dof)          # the case exists in this mesh
    tab = emf.build_chunk_tables(elem_dof, nd, 64)
    x = rng.normal(size=(nd, N))
    y, _ = emf.apply_operator(tab, emf.element_matrices(mesh, D), x, N)
    yref = (orc.stiffness_matrix(mesh, D, dof, nd) @ x.reshape(-1)).reshape(-1, N)
    assert rel_l2(y, yref) < 1e-13
