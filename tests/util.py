"""Shared helpers for the parity tests (oracle side only builds inputs / expected values)."""
import json
import os

import numpy as np

import meshfem_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

CANTILEVER_BC = {
    "no_rigid_motion": False,
    "regions": [
        {"type": "dirichlet", "value": [0, 0, 0],
         "box%": {"minCorner": [-0.0001, -0.0001, -0.0001], "maxCorner": [0.0001, 1.0001, 1.0001]}},
        {"type": "force", "value": [0, -10, 0],
         "box%": {"minCorner": [0.9999, -0.0001, -0.0001], "maxCorner": [1.0001, 1.0001, 1.0001]}},
    ],
}
CANTILEVER_2D_BC = {
    "no_rigid_motion": False,
    "regions": [
        {"type": "dirichlet", "value": [0, 0, 0],
         "box%": {"minCorner": [-0.0001, -0.0001, 0], "maxCorner": [0.0001, 1.0001, 0]}},
        {"type": "force", "value": [0, -10, 0],
         "box%": {"minCorner": [0.9999, -0.0001, 0], "maxCorner": [1.0001, 1.0001, 0]}},
    ],
}
ORTHO = {"type": "orthotropic", "young": [200, 120, 80], "poisson": [0.3, 0.2, 0.12, 0.3, 0.3, 0.18],
         "shear": [45, 35, 60]}


def grid_mesh(N, deg, sizes):
    V, T = orc.grid_simplices(list(sizes))
    return orc.build_mesh(N, deg, V, T)


def rel_l2(a, b):
    a = np.asarray(a).ravel(); b = np.asarray(b).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def cantilever_problem(N, deg, sizes, D=None):
    """Oracle-side set-up of the cantilever configs: mesh, material, fixed vars, load."""
    V, T = orc.grid_simplices(list(sizes))
    sim = orc.Simulator(N, deg, V, T)
    sim.set_material(orc.isotropic_D(N, 200.0, 0.35) if D is None else D)
    bc = CANTILEVER_BC if N == 3 else CANTILEVER_2D_BC
    conds, no_rigid, pps, pin = orc.read_boundary_conditions(N, bc, sim.mesh.bbox_min, sim.mesh.bbox_max)
    sim.apply_translation_pins(pin)
    sim.apply_boundary_conditions(conds)
    fixed, vals = sim.fixed_vars_and_values()
    f = sim.neumann_load()
    return sim, fixed, vals, f
