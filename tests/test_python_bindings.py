"""pybind11 modules python/{tensors,sparse_matrices,periodic_homogenization} -- the reference's Python
operator surface (src/python_bindings/*.cc) over this repository's host classes.  CPU tests cover what
needs no GPU (tensors, TripletMatrix); the GPU tests run SPSDSystem and homogenize/probe."""
import os
import sys

import numpy as np
import pytest

import meshfem_oracle as orc
from util import GOLDEN, ORTHO, ROOT, cantilever_problem, rel_l2


@pytest.fixture(scope="module")
def pymods(lib_built):
    sys.path.insert(0, os.path.join(ROOT, "python"))
    import periodic_homogenization
    import sparse_matrices
    import tensors
    return tensors, sparse_matrices, periodic_homogenization


def test_tensors_module(pymods, tmp_path):
    tensors = pymods[0]
    E = tensors.ElasticityTensor3D(200.0, 0.35)
    assert np.allclose(E.D, orc.isotropic_D(3, 200.0, 0.35), rtol=1e-15)
    assert abs(E.anisotropy() - 1) < 1e-12
    assert abs(E(0, 1, 0, 1) - E.D[5, 5]) == 0 and abs(E(0, 0, 1, 1) - E.D[0, 1]) == 0
    with pytest.raises(RuntimeError, match="Index out of bounds"):
        E(3, 0, 0, 0)
    E.setOrthotropic(200, 120, 80, 0.18, 0.12, 0.2, 45, 35, 60)
    D = orc.material_from_json(3, ORTHO)
    assert np.allclose(E.D, D, rtol=1e-13)
    assert np.allclose(E.getOrthotropicParameters(), [200, 120, 80, 0.18, 0.12, 0.2, 45, 35, 60], rtol=1e-12)
    s = np.array([0.1, -0.2, 0.3, 0.05, 0.07, -0.02])
    dbl = np.array([1, 1, 1, 2, 2, 2.0])
    assert np.allclose(E.doubleContract(s), D @ (dbl * s), rtol=1e-13)
    assert np.allclose(E.doubleContract(np.stack([s, 2 * s])), np.stack([D @ (dbl * s), 2 * D @ (dbl * s)]), rtol=1e-13)
    ed = E.computeEigenstrains()
    rt = np.diag(np.sqrt(dbl))
    assert np.allclose(ed.eigenvalues, np.linalg.eigvalsh(rt @ D @ rt), rtol=1e-12)
    k = 5      # E : s = lambda s for the returned eigenstrain (flattened, shear-doubling contraction)
    assert np.allclose(E.doubleContract(ed.eigenstrains[:, k]), ed.eigenvalues[k] * ed.eigenstrains[:, k], rtol=1e-9, atol=1e-9)
    assert np.allclose(E.inverse().D, np.linalg.inv(D) / np.outer(dbl, dbl), rtol=1e-12)
    assert (E - E).frobeniusNormSq() == 0
    # frobeniusNormSq = full rank-4 contraction (ElasticityTensor.hh:498-508); transform (:515-541)
    assert np.isclose(E.frobeniusNormSq(), orc.frobenius_norm_sq(3, D), rtol=1e-14)
    assert np.isclose(E.quadrupleContract(E.inverse()), (orc.tensor_C(3, D) * orc.tensor_C(3, E.inverse().D)).sum(), rtol=1e-12)
    R = np.array([[1.1, 0.2, 0.0], [0.05, 0.9, 0.1], [0.0, 0.15, 1.05]])
    assert np.allclose(E.transform(R).D, orc.transform_tensor(3, D, R), rtol=1e-13, atol=1e-12)
    th = 0.3
    Q = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    Eiso = tensors.ElasticityTensor3D(200.0, 0.35)
    assert np.allclose(Eiso.transform(Q).D, Eiso.D, atol=1e-12)          # isotropic tensors are rotation invariant
    assert np.isclose(E.transform(Q).frobeniusNormSq(), E.frobeniusNormSq(), rtol=1e-13)
    mat = tmp_path / "m.material"
    mat.write_text('{"type": "isotropic_material", "dim": 2, "young": 10.0, "poisson": 0.25}')
    E2 = tensors.ElasticityTensor2D(str(mat))
    assert np.allclose(E2.D, orc.isotropic_D(2, 10.0, 0.25), rtol=1e-14)
    E2.setIdentity()
    assert np.allclose(E2.D, np.diag([1, 1, 0.5]))
    assert "2D elasticity tensor" in repr(E2)
    assert np.allclose(tensors.ElasticityTensor3D(D).D, D)


def test_triplet_matrix(pymods, tmp_path):
    sm = pymods[1]
    A = sm.TripletMatrix(3, 3)
    A.addNZ(0, 0, 2.0); A.addNZ(0, 0, 1.0); A.addNZ(1, 2, 5.0); A.addNZ(2, 2, 0.0)
    assert A.nnz() == 4
    A.sumRepeated()                        # sums duplicates, prunes zeros (SparseMatrices.hh:280-374)
    assert A.nnz() == 2
    assert np.allclose(A.apply(np.ones(3)), [3, 5, 0])
    assert np.allclose(A.compressedColumn().toarray(), [[3, 0, 0], [0, 0, 5], [0, 0, 0]])
    p = str(tmp_path / "A.bin")
    A.dumpBinary(p)
    B = sm.TripletMatrix()
    B.readBinary(p)
    i, j, v = B.triplets()
    assert list(i) == [0, 1] and list(j) == [0, 2] and list(v) == [3.0, 5.0]


@pytest.mark.gpu
def test_spsd_system_solves_external_matrix(pymods):
    """sparse_matrices.SPSDSystem(K).fixVariables(...).solve(b) as in the reference's binding
    (sparse_matrices.cc:48-66), K = upper-triangle triplets assembled elsewhere (here: by the oracle)."""
    sm = pymods[1]
    sim, fixed, vals, f = cantilever_problem(3, 2, (5, 2, 2))
    I, J, W = orc.assemble_upper_triplets(sim.mesh, sim.D)
    K = sm.TripletMatrix(3 * sim.mesh.num_nodes, 3 * sim.mesh.num_nodes)
    K.addNZs(I, J, W)
    sys_ = sm.SPSDSystem(K)
    sys_.setTolerance(1e-12)
    sys_.fixVariables(list(fixed), list(vals))
    u = sys_.solve(f.reshape(-1))
    assert sys_.lastSolveInfo()["converged"]
    assert rel_l2(u, orc.solve_fixed(sim.stiffness(), f.reshape(-1), fixed, vals)) < 1e-8
    with pytest.raises(RuntimeError, match="Variable already fixed"):
        sys_.fixVariables([int(fixed[0])], [0.0])
    with pytest.raises(RuntimeError, match="Bad RHS"):
        sys_.solve(np.zeros(5))


@pytest.mark.gpu
@pytest.mark.parametrize("deg", [1, 2])
def test_homogenize_and_probe(pymods, deg):
    """periodic_homogenization.homogenize / probe (periodic_homogenization.cc:36-143) on the perforated
    golden cell: Ch and w_ij against the oracle, probe against its definition."""
    tensors, _, ph = pymods
    gold = np.load(os.path.join(GOLDEN, "homog_perforated.npz"))
    V, T = gold["V"], gold["T"]
    C = tensors.ElasticityTensor3D(200.0, 0.35)
    hr = ph.homogenize(V, T, C.D, degree=deg, centerFluctuationDisplacements=False, rtol=1e-12)
    assert np.abs(hr.Ch - gold[f"Eh_deg{deg}"]).max() < 1e-7 * np.abs(hr.Ch).max()
    for i in range(6):
        assert rel_l2(hr.w_ij[i], gold[f"w_deg{deg}"][i]) < 1e-6
    hc = ph.homogenize(V, T, C.D, degree=deg, rtol=1e-12)          # centred: every component averages to zero
    assert all(np.abs(w.mean(axis=0)).max() < 1e-12 for w in hc.w_ij)
    e = np.array([0.01, -0.02, 0.005, 0.003, 0.0, -0.004])
    u, strain = ph.probe(V, T, deg, hr, e)
    dbl = np.array([1, 1, 1, 2, 2, 2.0])
    w = sum(dbl[i] * e[i] * hr.w_ij[i] for i in range(6))
    m = orc.build_mesh(3, deg, V, T)
    Em = np.array([[e[0], e[5], e[4]], [e[5], e[1], e[3]], [e[4], e[3], e[2]]])
    bn = np.unique(m.bdry_elem_nodes)
    trans = np.array([w[bn][np.abs(m.nodes[bn][:, d] - m.bbox_min[d]) < 1e-9, d].mean() for d in range(3)])
    assert rel_l2(u, w - trans[None, :] + m.nodes @ Em.T) < 1e-12
    assert rel_l2(strain, sum(dbl[i] * e[i] * hr.strain_w_ij[i] for i in range(6)) + e[None, :]) < 1e-12
    # orthotropic base cell (positive octant): same tensor as the full periodic cell, w_ij as the oracle's
    from test_gpu_cli import _perforated_octant
    Vo, To = _perforated_octant()
    ho = ph.homogenize(Vo, To, C.D, degree=deg, orthotropicCell=True, centerFluctuationDisplacements=False, rtol=1e-12)
    assert np.abs(ho.Ch - gold[f"Eh_deg{deg}"]).max() < 1e-7 * np.abs(ho.Ch).max()
    sim = orc.Simulator(3, deg, Vo, To); sim.set_material(C.D)
    wo = orc.solve_orthotropic_cell_problems(sim)
    for i in range(6):
        assert rel_l2(ho.w_ij[i], wo[i]) < 1e-6
    # identified node pairs from a file (PeriodicCondition(mesh, pcFile), BoundaryConditions.hh:563-610): the pairs the
    # matcher finds, written out, reproduce the matcher's tensor; a missing file raises the reference's message
    import tempfile
    dof = np.asarray(orc.periodic_condition(m)[0])
    order = np.argsort(dof, kind="stable")
    same = dof[order][1:] == dof[order][:-1]
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "pairs.txt")
        with open(path, "w") as fh:
            fh.write("".join(f"{a} {b}\n" for a, b in zip(order[:-1][same], order[1:][same])))
        hm = ph.homogenize(V, T, C.D, degree=deg, manualPeriodicVerticesFile=path, rtol=1e-12)
    assert np.abs(hm.Ch - gold[f"Eh_deg{deg}"]).max() < 1e-7 * np.abs(hm.Ch).max()
    with pytest.raises(RuntimeError, match="Couldn't open"):
        ph.homogenize(V, T, C.D, degree=deg, manualPeriodicVerticesFile="/nonexistent/x.txt")
