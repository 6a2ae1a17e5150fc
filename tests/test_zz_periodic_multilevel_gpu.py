"""GPU: the multilevel preconditioner on a PERIODIC cell problem large enough for the automatic rule (>= 30k DoFs):
periodic DoFs (several nodes per DoF, couplings that wrap around the cell), a pinned variable, six right-hand sides
solved one after the other -- against the batched block-Jacobi PCG of the same handle configuration
(coarse_aggregates = 0), which the small-size tests pin to the oracle (tests/test_gpu_cli.py, homog_perforated golden).
Also the theory KAT at this size: the homogenized tensor of the perforated cubic cell is cubic-symmetric."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_periodic_cell_problems_multilevel_equals_block_jacobi(lib_built):
    from meshfem_b200 import distributed, hostlib
    E, nu = 200.0, 0.35
    lam, mu = nu * E / ((1 + nu) * (1 - 2 * nu)), E / (2 + 2 * nu)
    D = np.zeros((6, 6)); D[:3, :3] = lam; D[np.arange(3), np.arange(3)] = lam + 2 * mu; D[np.arange(3, 6), np.arange(3, 6)] = mu
    raw = hostlib.perforated_cell(3, 12, 6)                  # 12^3 voxels minus 6^3: 36,288 quadratic tets, ~56k nodes
    Eh0, x0 = distributed.homogenize(raw, 2, D, rtol=1e-11, return_fields=True, coarse_aggregates=0)
    Eh1, x1 = distributed.homogenize(raw, 2, D, rtol=1e-11, return_fields=True, coarse_aggregates=-1)
    Eh2, x2 = distributed.homogenize(raw, 2, D, rtol=1e-11, return_fields=True, coarse_aggregates=64, coarse_fine_nodes=24)
    it0 = max(s["iterations"] for s in x0["solves"])
    for Eh, x in ((Eh1, x1), (Eh2, x2)):
        assert np.abs(Eh - Eh0).max() < 1e-9 * np.abs(Eh0).max()
        for w, w0 in zip(x["w"], x0["w"]):
            assert np.linalg.norm(w - w0) < 1e-7 * np.linalg.norm(w0)
        assert all(s["converged"] for s in x["solves"])
        assert max(s["iterations"] for s in x["solves"]) < 0.7 * it0, ([s["iterations"] for s in x["solves"]], it0)
    assert np.abs(Eh1 - Eh1.T).max() < 1e-10 * np.abs(Eh1).max()
    assert max(abs(Eh1[0, 0] - Eh1[1, 1]), abs(Eh1[0, 0] - Eh1[2, 2]), abs(Eh1[3, 3] - Eh1[4, 4])) < 1e-8 * abs(Eh1[0, 0])
