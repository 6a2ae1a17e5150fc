"""Generates tests/golden/microstructures.npz: the reference's own example microstructures
(/root/reference/examples/meshes/{2D_microstructure, 2D_microstructure_orthocell, 3D_microstructure_orthocell}.msh,
the inputs of python/examples/Homogenization.ipynb) as vertex / simplex arrays, with the homogenized elasticity
tensors the CPU oracle computes for them with the notebook's base material (E = 200, nu = 0.35): full periodic
homogenization for the 2D cell, orthotropic-base-cell homogenization for the two positive-orthant cells.
The notebook's own self-check -- the orthotropic-cell route gives the tensor of the full periodic cell -- holds for the
oracle on the 2D pair to 1e-13 (asserted here and in tests/test_oracle_kats.py).  Run in the build container (needs
/root/reference, which does not exist on the GPU box): python tests/golden/make_microstructure_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import meshfem_oracle as orc  # noqa: E402
from meshfem_b200 import hostlib  # noqa: E402  (MeshIO: binary Gmsh 2.2 reader)

MESHES = "/root/reference/examples/meshes/"


def load(name, dim):
    V, T = hostlib.load_mesh(MESHES + name).arrays()
    return np.ascontiguousarray(V[:, :dim]), np.ascontiguousarray(T, dtype=np.int32)


def main():
    out = {}
    D2, D3 = orc.isotropic_D(2, 200.0, 0.35), orc.isotropic_D(3, 200.0, 0.35)
    V, T = load("2D_microstructure.msh", 2)
    out["V_2d_full"], out["T_2d_full"] = V, T
    for deg in (1, 2):
        sim = orc.Simulator(2, deg, V, T); sim.set_material(D2)
        out[f"Eh_2d_full_deg{deg}"] = orc.homogenized_tensor_displacement_form(sim, orc.solve_cell_problems(sim))
    V, T = load("2D_microstructure_orthocell.msh", 2)
    out["V_2d_ortho"], out["T_2d_ortho"] = V, T
    for deg in (1, 2):
        sim = orc.Simulator(2, deg, V, T); sim.set_material(D2)
        Eh = orc.orthotropic_homogenized_tensor_displacement_form(sim, orc.solve_orthotropic_cell_problems(sim))
        full = out[f"Eh_2d_full_deg{deg}"]
        assert np.abs(Eh - full).max() < 1e-11 * np.abs(full).max()
        out[f"Eh_2d_ortho_deg{deg}"] = Eh
    V, T = load("3D_microstructure_orthocell.msh", 3)
    out["V_3d_ortho"], out["T_3d_ortho"] = V, T
    for deg in (1, 2):
        sim = orc.Simulator(3, deg, V, T); sim.set_material(D3)
        out[f"Eh_3d_ortho_deg{deg}"] = orc.orthotropic_homogenized_tensor_displacement_form(sim, orc.solve_orthotropic_cell_problems(sim))
    # the FULL 3D cell (68,888 vertices; its 7 MB mesh is not committed): periodic homogenization, degree 1, takes the
    # oracle ~95 s -- computed with --full3d, otherwise carried over from the existing file.  It agrees with the
    # orthotropic-cell tensor to 5.5e-13 (the notebook's "Moduli discrepancy" check on the reference's real 3D pair).
    path = os.path.join(HERE, "microstructures.npz")
    if "--full3d" in sys.argv:
        Vf, Tf = load("3D_microstructure.msh", 3)
        sim = orc.Simulator(3, 1, Vf, Tf); sim.set_material(D3)
        out["Eh_3d_full_deg1"] = orc.homogenized_tensor_displacement_form(sim, orc.solve_cell_problems(sim))
    elif os.path.exists(path) and "Eh_3d_full_deg1" in np.load(path).files:
        out["Eh_3d_full_deg1"] = np.load(path)["Eh_3d_full_deg1"]
    if "Eh_3d_full_deg1" in out:
        d = np.abs(out["Eh_3d_full_deg1"] - out["Eh_3d_ortho_deg1"]).max() / np.abs(out["Eh_3d_ortho_deg1"]).max()
        print("3D full periodic vs orthotropic cell, degree 1: relative discrepancy", d)
        assert d < 1e-10
    np.savez_compressed(path, **out)
    for k, v in out.items():
        print(k, v.shape)


if __name__ == "__main__":
    main()
