"""GPU: shape sensitivity of the homogenized tensor through the Python binding (cell problems and the
delta-fluctuation solves on the device, the discrete differential on the host) against the oracle, which is itself
pinned by finite differences (tests/test_shape_derivatives.py)."""
import os
import sys

import numpy as np
import pytest

import meshfem_oracle as orc
from util import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,deg", [(2, 2), (3, 1), (3, 2)])
def test_binding_shape_derivative_matches_oracle(lib_built, N, deg):
    sys.path.insert(0, os.path.join(ROOT, "python"))
    import periodic_homogenization as ph
    from meshfem_b200 import hostlib
    raw = hostlib.perforated_cell(N, 4, 2)
    V, T = raw.arrays()
    D = orc.isotropic_D(N, 200.0, 0.35)
    sim = orc.Simulator(N, deg, V, T)
    sim.set_material(D)
    w = orc.solve_cell_problems(sim)
    rng = np.random.default_rng(5)
    inner = ((V[:, :N] > 1e-9) & (V[:, :N] < 1 - 1e-9)).all(axis=1)
    dp = rng.standard_normal((V.shape[0], N)) * inner[:, None] * 0.1
    r = ph.shapeDerivative(V, T, D, degree=deg, deltaP=dp, rtol=1e-12)
    Eh = orc.homogenized_tensor_displacement_form(sim, w)
    assert np.abs(r["Ch"] - Eh).max() < 1e-8 * np.abs(Eh).max()
    dCh = orc.homogenized_tensor_discrete_differential(sim, w)
    assert np.abs(r["dCh"] - dCh).max() < 1e-7 * np.abs(dCh).max()
    dw = np.array(orc.delta_fluctuation_displacements(sim, w, dp))
    assert np.abs(r["delta_w_ij"] - dw).max() < 1e-6 * np.abs(dw).max()
    assert np.abs(r["delta_Ch"] - np.einsum("vcfg,vc->fg", dCh, dp)).max() < 1e-7 * np.abs(dCh).max()
