"""Synthetic workloads of BASELINE.json (host-side input generation for bench.py and tests).

The cantilever boundary conditions are examples/cantilever/cantilever.bc of the reference
applied with the reference's semantics (LinearElasticity.hh:881-1027): Dirichlet on boundary
NODES inside the box, `force` spread over boundary ELEMENTS selected by the barycentre of
their vertices and divided by the region area; nodal loads by integrated shape functions
(LinearElasticity.hh:341-347, 703-717).  Vectorised numpy; checked against the oracle's
general implementation in tests/test_workloads.py.
"""
from __future__ import annotations

import numpy as np

from meshfem_b200 import hostlib

CONFIGS = {
    # name: (grid, degree, material)  -- BASELINE.md section 4
    "cfg2": ((100, 20, 20), 1, "iso"),
    "cfg3": ((130, 26, 26), 2, "ortho"),
    "cfg5": ((220, 44, 44), 2, "iso"),
}


def isotropic_D3(E=200.0, nu=0.35):
    lam = nu * E / ((1 + nu) * (1 - 2 * nu)); mu = E / (2 + 2 * nu)
    D = np.zeros((6, 6)); D[:3, :3] = lam
    D[np.arange(3), np.arange(3)] = lam + 2 * mu
    D[np.arange(3, 6), np.arange(3, 6)] = mu
    return D


def orthotropic_D3():
    """BASELINE.md cfg3 material: young [200,120,80], poisson [0.3,0.2,0.12,0.3,0.3,0.18]
    (yz,zy,zx,xz,xy,yx), shear [45,35,60] (yz,zx,xy); ElasticityTensor.hh:136-152."""
    Ex, Ey, Ez = 200.0, 120.0, 80.0
    nu_yz, nu_zy, nu_zx, nu_xz, nu_xy, nu_yx = 0.3, 0.2, 0.12, 0.3, 0.3, 0.18
    M = np.zeros((6, 6))
    M[0, 0] = 1 / Ex; M[0, 1] = -nu_yx / Ey; M[0, 2] = -nu_zx / Ez
    M[1, 1] = 1 / Ey; M[1, 2] = -nu_zy / Ez; M[2, 2] = 1 / Ez
    M[3, 3] = 1 / 45.0; M[4, 4] = 1 / 35.0; M[5, 5] = 1 / 60.0
    M = np.triu(M) + np.triu(M, 1).T
    return np.linalg.inv(M)


def material(name):
    return isotropic_D3() if name == "iso" else orthotropic_D3()


def grid_femmesh(grid, deg):
    return hostlib.grid(list(grid)).femmesh(deg)


def cantilever_inputs(m, force=(0.0, -10.0, 0.0)):
    """(fixed_vars, fixed_vals, f[nNodes, N]) for examples/cantilever/cantilever.bc on mesh m."""
    N = m.N
    lo, hi = m.bbox_min, m.bbox_max
    ext = hi - lo
    # dirichlet region: box% [-1e-4, 1e-4] x [-1e-4, 1+1e-4]^(N-1)
    dmin = lo + np.array([-0.0001] * N) * ext
    dmax = lo + np.array([0.0001] + [1.0001] * (N - 1)) * ext
    P = m.nodes[m.bdry_nodes]
    inside = np.all((P >= dmin) & (P <= dmax), axis=1)
    dn = m.bdry_nodes[inside].astype(np.int64)
    fixed = (N * dn[:, None] + np.arange(N)[None, :]).reshape(-1)
    vals = np.zeros(fixed.size)
    # force region: box% [0.9999, 1.0001] x [-1e-4, 1+1e-4]^(N-1), boundary elements by vertex barycentre
    fmin = lo + np.array([0.9999] + [-0.0001] * (N - 1)) * ext
    fmax = lo + np.array([1.0001] * N) * ext
    centers = m.nodes[m.bdry_elem_vertices].mean(axis=1)
    sel = np.all((centers >= fmin) & (centers <= fmax), axis=1)
    if not sel.any():
        raise RuntimeError("Neumann region unmatched")
    area = m.bdry_vol[sel].sum()
    traction = np.asarray(force[:N]) / area
    if m.deg == 1:
        w = np.full(N, 1.0 / N)
    elif N == 3:
        w = np.array([0, 0, 0, 1 / 3, 1 / 3, 1 / 3])
    else:
        w = np.array([1 / 6, 1 / 6, 4 / 6])
    f = np.zeros((m.num_nodes, N))
    bn = m.bdry_elem_nodes[sel]
    contrib = (m.bdry_vol[sel][:, None] * w[None, :])[:, :, None] * traction[None, None, :]
    np.add.at(f, bn.reshape(-1), contrib.reshape(-1, N))
    return fixed, vals, f
