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


B9CREATOR = {"type": "isotropic_material", "dim": 3, "density": 1.0, "young": 200.0, "poisson": 0.35}


def read_msh_fields(path):
    """Fields written by MSHFieldWriter ($NodeData / $ElementData, binary or ascii) -> {name: array}."""
    import struct
    data = open(path, "rb").read()
    binary = data.split(b"\n", 2)[1].split()[1] == b"1"
    out = {}
    pos = 0
    while True:
        i1 = data.find(b"$NodeData\n", pos)
        i2 = data.find(b"$ElementData\n", pos)
        i3 = data.find(b"$ElementNodeData\n", pos)
        cands = [i for i in (i1, i2, i3) if i >= 0]
        if not cands:
            break
        i = min(cands)
        elem_node = i == i3
        p = data.index(b"\n", i) + 1
        def line():
            nonlocal p
            e = data.index(b"\n", p); s = data[p:e].decode(); p = e + 1
            return s
        assert line() == "1"
        name = line().strip('"')
        assert line() == "0" and line() == "3" and line() == "0"
        dim = int(line()); n = int(line())
        if elem_node:        # per element: id, nodes per element, then nodesPerElem x dim values -> (n, npe, dim)
            if binary:
                npe = int(np.frombuffer(data, dtype="<i4", count=2, offset=p)[1])
                rec = np.frombuffer(data, dtype=np.dtype([("id", "<i4"), ("npe", "<i4"), ("v", "<f8", npe * dim)]), count=n, offset=p)
                assert np.array_equal(rec["id"], np.arange(1, n + 1)) and (rec["npe"] == npe).all()
                out[name] = rec["v"].reshape(n, npe, dim).copy(); p += n * (8 + 8 * npe * dim)
            else:
                rows = [line().split() for _ in range(n)]
                npe = int(rows[0][1])
                out[name] = np.array([[float(x) for x in r[2:]] for r in rows]).reshape(n, npe, dim)
            pos = p
            continue
        if binary:
            rec = np.frombuffer(data, dtype=np.dtype([("id", "<i4"), ("v", "<f8", dim)]), count=n, offset=p)
            assert np.array_equal(rec["id"], np.arange(1, n + 1))
            out[name] = rec["v"].reshape(n, dim).copy(); p += n * (4 + 8 * dim)
        else:
            vals = np.zeros((n, dim))
            for k in range(n):
                t = line().split(); vals[k] = [float(x) for x in t[1:]]
            out[name] = vals
        pos = p
    return out


def sym9_to_flat(N, a9):
    """9 row-major doubles (padded 3x3) -> flattened Voigt (n, flat)."""
    idx = [(0, 0), (1, 1), (0, 1)] if N == 2 else [(0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1)]
    return np.stack([a9[:, 3 * i + j] for i, j in idx], axis=1)
