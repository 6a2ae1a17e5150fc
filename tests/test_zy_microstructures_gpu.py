"""GPU: the reference's example microstructures (tests/golden/microstructures.npz: examples/meshes/2D_microstructure,
2D_microstructure_orthocell, 3D_microstructure_orthocell, the inputs of python/examples/Homogenization.ipynb) through
the Python binding, as the notebook does it -- periodic homogenization of the full 2D cell, orthotropic base-cell
homogenization of the positive-orthant cells -- against the oracle's tensors (1e-7 relative to max |Eh|), and the
notebook's closing check: both routes give the same tensor.  Written after the round-1 GPU budget was spent."""
import os
import sys

import numpy as np
import pytest

from util import ROOT

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden", "microstructures.npz")


@pytest.fixture(scope="module")
def ph(lib_built):
    sys.path.insert(0, os.path.join(ROOT, "python"))
    import periodic_homogenization
    return periodic_homogenization


def _iso(N, E=200.0, nu=0.35):
    import meshfem_oracle as orc
    return orc.isotropic_D(N, E, nu)


@pytest.mark.parametrize("deg", [1, 2])
def test_2d_microstructure_periodic_and_orthotropic_cell(ph, deg):
    g = np.load(GOLD)
    full = ph.homogenize(g["V_2d_full"], g["T_2d_full"], _iso(2), degree=deg, rtol=1e-12)
    ortho = ph.homogenize(g["V_2d_ortho"], g["T_2d_ortho"], _iso(2), degree=deg, orthotropicCell=True, rtol=1e-12)
    want = g[f"Eh_2d_full_deg{deg}"]
    scale = np.abs(want).max()
    assert np.abs(full.Ch - want).max() < 1e-7 * scale
    assert np.abs(ortho.Ch - g[f"Eh_2d_ortho_deg{deg}"]).max() < 1e-7 * scale
    assert np.abs(ortho.Ch - full.Ch).max() < 1e-7 * scale          # "Moduli discrepancy" of the notebook


@pytest.mark.parametrize("deg", [1, 2])
def test_3d_microstructure_orthotropic_cell(ph, deg):
    g = np.load(GOLD)
    hr = ph.homogenize(g["V_3d_ortho"], g["T_3d_ortho"], _iso(3), degree=deg, orthotropicCell=True, rtol=1e-12)
    want = g[f"Eh_3d_ortho_deg{deg}"]
    assert np.abs(hr.Ch - want).max() < 1e-7 * np.abs(want).max()
