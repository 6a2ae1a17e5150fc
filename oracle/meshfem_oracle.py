"""CPU oracle for the MeshFEM linear-elasticity assemble-and-solve path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (meshfem_b200/, include/,
the C-ABI library, the host C++ surface) may import, link or execute this file.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs use it, and only as the checker.

This is a from-scratch numpy restatement of the reference algorithm; every
function cites the reference file:line it follows (paths relative to
/root/reference/src/lib/MeshFEM unless noted).

PINNING STATUS: the reference cannot be compiled in this image (it needs Eigen,
SuiteSparse, TBB, Boost, nlohmann/json, tinyexpr -- none present, no network),
and its own test-suite holds no golden vector for perElementStiffness, the
assembled system, or a solve.  The element-math layer of this oracle IS pinned
against the reference's own known-answer tests (quadrature monomial tables,
shape-function identities, material JSON fixtures, flatten/unflatten tables:
tests/test_oracle_kats.py); the assemble/solve boundary is "parity unpinned"
by the reference and is pinned instead by theory KATs (rigid-body null space,
patch test, solid-cell homogenisation == base tensor, symmetric positive
definiteness, three independent formulations of Ke agreeing to 1e-15).
The direct solve stands in for CHOLMOD with SuperLU (scipy.sparse.linalg.splu):
the SPD solution is unique, so any backward-stable direct solver is an equally
valid reference at the 1e-6 tolerance.
"""
from __future__ import annotations

import json
import math
import struct
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# ----------------------------------------------------------------------------
# Simplex tables (Simplex.hh:16-46)
# ----------------------------------------------------------------------------
EDGE_START = (0, 1, 2, 0, 2, 1)   # Simplex.hh:39-41
EDGE_END = (1, 2, 0, 3, 3, 3)


def num_vertices(K):            # Simplex.hh:16
    return K + 1


def num_edges(K):               # Simplex.hh:17
    return K * (K + 1) // 2


def num_nodes(K, deg):          # Simplex.hh:22-27
    if K == 1:
        return deg + 1
    if K == 2:
        return (deg + 1) * (deg + 2) // 2
    if K == 3:
        return (deg + 1) * (deg + 2) * (deg + 3) // 6
    raise ValueError("Simplex dimension must be 1, 2, or 3")


# ----------------------------------------------------------------------------
# Tensor flattening (Flattening.hh:21-83)
# ----------------------------------------------------------------------------
def flat_len(dim):              # Flattening.hh:21
    return dim * (dim + 1) // 2


def flatten_indices(dim, i, j):  # Flattening.hh:47-60
    if i == j:
        return i
    if dim == 2:
        return 2
    lo, hi = (i, j) if i < j else (j, i)
    return 4 - lo if hi == 2 else 5


def unflatten_index(dim, i):    # Flattening.hh:64-83
    if dim == 2:
        return (i, i) if i < 2 else (0, 1)
    if i < 3:
        return (i, i)
    return {3: (1, 2), 4: (0, 2), 5: (0, 1)}[i]


# ----------------------------------------------------------------------------
# Elasticity tensors (ElasticityTensor.hh:100-164, 274-289, 315-323, 435-447)
# ----------------------------------------------------------------------------
def isotropic_D(dim, E, nu):
    """ElasticityTensor::setIsotropic (ElasticityTensor.hh:100-134)."""
    lam = (nu * E) / ((1.0 + nu) * (1.0 - 2.0 * nu))
    mu = E / (2.0 + 2.0 * nu)
    if dim == 2:                      # plane stress, :111-112
        lam = (nu * E) / (1.0 - nu * nu)
    n = flat_len(dim)
    D = np.zeros((n, n))
    D[:dim, :dim] = lam
    for i in range(dim):
        D[i, i] = lam + 2 * mu
    for i in range(dim, n):
        D[i, i] = mu
    return D


def orthotropic_D3(Ex, Ey, Ez, nuYX, nuZX, nuZY, muYZ, muZX, muXY):
    """setOrthotropic3D (ElasticityTensor.hh:136-152): invert the symmetrised
    upper triangle of the compliance-like matrix."""
    M = np.zeros((6, 6))
    M[0, 0] = 1.0 / Ex; M[0, 1] = -nuYX / Ey; M[0, 2] = -nuZX / Ez
    M[1, 1] = 1.0 / Ey; M[1, 2] = -nuZY / Ez
    M[2, 2] = 1.0 / Ez
    M[3, 3] = 1.0 / muYZ; M[4, 4] = 1.0 / muZX; M[5, 5] = 1.0 / muXY
    M = np.triu(M) + np.triu(M, 1).T
    return _upper(np.linalg.inv(M))


def _upper(D):
    """A major-symmetric ElasticityTensor IS its upper triangle: D(i,j) reads m_d(min,max)
    (ElasticityTensor.hh:279-283)."""
    return np.triu(D) + np.triu(D, 1).T


def orthotropic_D2(Ex, Ey, nuYX, muXY):
    """setOrthotropic2D (ElasticityTensor.hh:154-164)."""
    M = np.zeros((3, 3))
    M[0, 0] = 1.0 / Ex; M[0, 1] = -nuYX / Ey
    M[1, 1] = 1.0 / Ey
    M[2, 2] = 1.0 / muXY
    M = np.triu(M) + np.triu(M, 1).T
    return _upper(np.linalg.inv(M))


def material_from_json(dim, cfg):
    """Materials::Constant<N>::setFromJson (Materials.cc:194-300)."""
    t = cfg["type"]
    if t in ("isotropic_material", "isotropic"):
        return isotropic_D(dim, float(cfg["young"]), float(cfg["poisson"]))
    if t in ("orthotropic_material", "orthotropic"):
        young = [float(v) for v in cfg["young"]]
        poisson = [float(v) for v in cfg["poisson"]]
        shear = [float(v) for v in cfg["shear"]]
        if dim == 2:
            if (len(young), len(poisson), len(shear)) != (2, 2, 1):
                raise RuntimeError("Invalid orthotropic parameter vector size")
            Ex, Ey = young
            nu_xy, nu_yx = poisson
            D = orthotropic_D2(Ex, Ey, nu_yx, shear[0])
            if abs(nu_yx / Ey - nu_xy / Ex) > 1e-10:
                raise RuntimeError("Orthotopic parameters violate symmetry")
            return D
        if (len(young), len(poisson), len(shear)) != (3, 6, 3):
            raise RuntimeError("Invalid orthotropic parameter vector size")
        Ex, Ey, Ez = young
        nu_yz, nu_zy, nu_zx, nu_xz, nu_xy, nu_yx = poisson
        mu_yz, mu_zx, mu_xy = shear
        D = orthotropic_D3(Ex, Ey, Ez, nu_yx, nu_zx, nu_zy, mu_yz, mu_zx, mu_xy)
        if (abs(nu_yx / Ey - nu_xy / Ex) > 1e-10 or abs(nu_yz / Ey - nu_zy / Ez) > 1e-10
                or abs(nu_zx / Ez - nu_xz / Ex) > 1e-10):
            raise RuntimeError("Orthotopic parameters violate symmetry")
        return D
    if t in ("symmetric_material", "anisotropic"):
        n = flat_len(dim)
        D = np.zeros((n, n))
        rows = cfg["material_matrix"]
        for r, row in enumerate(rows):
            if len(row) != n:
                raise RuntimeError("Failed to parse material_matrix")
            for c, val in enumerate(row):
                if r <= c:
                    D[r, c] = D[c, r] = float(val)
                elif abs(D[r, c] - float(val)) > 1e-10:
                    raise RuntimeError("Asymmetric material_matrix")
        return D
    raise RuntimeError("Invalid type.")


def material_from_file(dim, path):
    with open(path) as f:
        return material_from_json(dim, json.load(f))


def tensor_C(dim, D):
    """Rank-4 view C_ijkl = D(flat(i,j), flat(k,l)) (ElasticityTensor.hh:274-277)."""
    C = np.zeros((dim,) * 4)
    for i in range(dim):
        for j in range(dim):
            for k in range(dim):
                for l in range(dim):
                    C[i, j, k, l] = D[flatten_indices(dim, i, j), flatten_indices(dim, k, l)]
    return C


def double_contract(dim, D, eps_flat):
    """D * shearDoubled(eps) (ElasticityTensor.hh:435-447)."""
    e = np.array(eps_flat, dtype=float, copy=True)
    e[..., dim:] *= 2.0
    return e @ D.T


def tensor_inverse(dim, D):
    """ElasticityTensor::inverse (ElasticityTensor.hh:315-323): invert flattened
    matrix, then halve shear rows and columns."""
    S = np.linalg.inv(D)
    S[dim:, :] *= 0.5
    S[:, dim:] *= 0.5
    return S


# ----------------------------------------------------------------------------
# Quadrature (GaussQuadrature.hh:115-127 tri, :283-295 tet) and shape functions
# (Functions.hh, EmbeddedElement.hh:288-332)
# ----------------------------------------------------------------------------
TET_C0 = 0.58541019662496845446
TET_C1 = 0.13819660112501051518


def quadrature_points(K, deg):
    """Barycentric points and weights (summing to 1) of Quadrature<K,deg>.
    deg 0/1: centroid rule; deg 2: tri 3-pt (GaussQuadrature.hh:115-127),
    tet 4-pt (:283-295); edge deg 2: 2-pt Gauss."""
    if deg <= 1:
        return np.full((1, K + 1), 1.0 / (K + 1)), np.array([1.0])
    if deg == 2:
        if K == 3:
            P = np.full((4, 4), TET_C1)
            np.fill_diagonal(P, TET_C0)
            return P, np.full(4, 0.25)
        if K == 2:
            P = np.full((3, 3), 1.0 / 6.0)
            np.fill_diagonal(P, 2.0 / 3.0)
            return P, np.full(3, 1.0 / 3.0)
        if K == 1:
            a = 0.5 + 0.5 / math.sqrt(3.0)
            return np.array([[a, 1 - a], [1 - a, a]]), np.array([0.5, 0.5])
    raise NotImplementedError("oracle carries only the rules the hot path uses")


def shape_functions(K, deg, lam):
    """phi_i(lambda) for deg 1/2 (Functions.hh:98-150)."""
    lam = np.asarray(lam)
    if deg == 1:
        return lam.copy()
    out = np.zeros(lam.shape[:-1] + (num_nodes(K, 2),))
    for i in range(K + 1):
        out[..., i] = lam[..., i] * (2 * lam[..., i] - 1)
    for k in range(num_edges(K)):
        out[..., K + 1 + k] = 4 * lam[..., EDGE_START[k]] * lam[..., EDGE_END[k]]
    return out


def grad_phi_coeffs(K, deg, lam):
    """alpha[i][a]: grad phi_i(lam) = sum_a alpha[i][a] grad lambda_a
    (pointwise form of EmbeddedElement.hh:315-332)."""
    nn = num_nodes(K, deg)
    A = np.zeros((nn, K + 1))
    if deg == 1:
        A[:, :] = np.eye(K + 1)
        return A
    for i in range(K + 1):
        A[i, i] = 4 * lam[i] - 1
    for k in range(num_edges(K)):
        s, e = EDGE_START[k], EDGE_END[k]
        A[K + 1 + k, s] = 4 * lam[e]
        A[K + 1 + k, e] = 4 * lam[s]
    return A


def grad_phi_interpolant(K, deg):
    """Interpolant form (EmbeddedElement.hh:288-313): T[i][v][a] such that the
    deg-1-lower interpolant of grad phi_i has nodal value sum_a T[i][v][a] grad
    lambda_a at vertex v (deg 2), or the single constant value (deg 1, v = 0)."""
    if deg == 1:
        T = np.zeros((K + 1, 1, K + 1))
        for i in range(K + 1):
            T[i, 0, i] = 1.0
        return T
    nn = num_nodes(K, 2)
    T = np.zeros((nn, K + 1, K + 1))
    for i in range(K + 1):
        for v in range(K + 1):
            T[i, v, i] = 3.0 if v == i else -1.0       # :296-300
    for k in range(num_edges(K)):
        s, e = EDGE_START[k], EDGE_END[k]
        T[K + 1 + k, s, e] = 4.0                        # :303-310
        T[K + 1 + k, e, s] = 4.0
    return T


def integrated_phis(K, deg):
    """int phi_i / vol (Functions.hh:247-274): deg1 1/(K+1); deg2: K=1 (1/6,1/6,4/6),
    K=2 vertices 0, edges 1/3; K=3 vertices -1/20, edges 1/5."""
    if deg == 1:
        return np.full(K + 1, 1.0 / (K + 1))
    if K == 1:
        return np.array([1 / 6, 1 / 6, 4 / 6])
    if K == 2:
        return np.array([0, 0, 0, 1 / 3, 1 / 3, 1 / 3])
    return np.array([-1 / 20] * 4 + [1 / 5] * 6)


# ----------------------------------------------------------------------------
# Embedding (EmbeddedElement.hh:87-104, 128-149, 170-190, 211-231)
# ----------------------------------------------------------------------------
def embed_simplices(P):
    """P: (ne, K+1, N) vertex coordinates of full-dimensional simplices (K == N).
    Returns vol (ne,), G (ne, N, K+1) with G[:, :, k] = grad lambda_k."""
    P = np.asarray(P, dtype=float)
    K = P.shape[1] - 1
    if K == 3:                                   # EmbeddedElement.hh:211-231
        p0, p1, p2, p3 = (P[:, i] for i in range(4))
        n0 = np.cross(p3 - p1, p2 - p1)
        V6 = np.einsum("ij,ij->i", p0 - p1, n0)
        G = np.empty((P.shape[0], 3, 4))
        G[:, :, 0] = n0
        G[:, :, 1] = np.cross(p2 - p0, p3 - p0)
        G[:, :, 2] = np.cross(p3 - p0, p1 - p0)
        G[:, :, 3] = np.cross(p1 - p0, p2 - p0)
        G /= V6[:, None, None]
        return V6 / 6.0, G
    if K == 2:                                   # EmbeddedElement.hh:170-190
        p0, p1, p2 = (P[:, i] for i in range(3))
        e = [p2 - p1, p0 - p2, p1 - p0]
        dblA = e[1][:, 0] * e[2][:, 1] - e[1][:, 1] * e[2][:, 0]
        G = np.empty((P.shape[0], 2, 3))
        for k in range(3):
            G[:, 0, k] = -e[k][:, 1]
            G[:, 1, k] = e[k][:, 0]
        G /= dblA[:, None, None]
        return dblA / 2.0, G
    raise ValueError


def embed_boundary(P):
    """Boundary simplices: (nb, K, N) with K = N. Returns volume (length/area)
    and outward unit normal (EmbeddedElement.hh:87-104 edge-in-2D, :128-149
    tri-in-3D)."""
    P = np.asarray(P, dtype=float)
    N = P.shape[2]
    if N == 2:
        e = P[:, 1] - P[:, 0]
        L = np.linalg.norm(e, axis=1)
        n = np.stack([-e[:, 1], e[:, 0]], axis=1) / L[:, None]
        return L, n
    e1 = P[:, 0] - P[:, 2]
    e2 = P[:, 1] - P[:, 0]
    n = np.cross(e1, e2)
    dblA = np.linalg.norm(n, axis=1)
    return dblA / 2.0, n / dblA[:, None]


# ----------------------------------------------------------------------------
# Synthetic meshes: `grid -t` (src/bin/tools/grid.cc:115-137,
# filters/gen_grid.hh:14-92, hex_tet_subdiv.hh:32-104, quad_tri_subdiv.hh)
# ----------------------------------------------------------------------------
HEX_FACES = ((0, 3, 2, 1), (0, 4, 7, 3), (4, 5, 6, 7), (1, 2, 6, 5), (0, 1, 5, 4), (2, 3, 7, 6))


def gen_grid(sizes):
    """gen_grid (filters/gen_grid.hh:14-92): integer-lattice vertices, quads/hexes
    in Gmsh order."""
    if len(sizes) == 2:
        nC, nR = sizes
        V = np.array([(c, r, 0.0) for r in range(nR + 1) for c in range(nC + 1)], dtype=float)
        idx = lambda r, c: (nC + 1) * r + c
        E = np.array([(idx(r, c), idx(r, c + 1), idx(r + 1, c + 1), idx(r + 1, c))
                      for r in range(nR) for c in range(nC)], dtype=np.int64)
        return V, E
    nC, nR, nS = sizes
    V = np.array([(c, r, s) for s in range(nS + 1) for r in range(nR + 1) for c in range(nC + 1)],
                 dtype=float)
    idx = lambda s, r, c: (nC + 1) * ((nR + 1) * s + r) + c
    E = np.array([(idx(s, r, c), idx(s, r, c + 1), idx(s, r + 1, c + 1), idx(s, r + 1, c),
                   idx(s + 1, r, c), idx(s + 1, r, c + 1), idx(s + 1, r + 1, c + 1), idx(s + 1, r + 1, c))
                  for s in range(nS) for r in range(nR) for c in range(nC)], dtype=np.int64)
    return V, E


def hex_tet_subdiv(V, H):
    """hex_tet_subdiv (filters/hex_tet_subdiv.hh:32-104): 24 tets per hex."""
    outV = [tuple(v) for v in V]
    tets = []
    face_center = {}
    for e in H:
        hc = len(outV)
        outV.append(tuple(np.sum(V[e], axis=0) / 8))
        for f in HEX_FACES:
            key = tuple(sorted(int(e[c]) for c in f))
            fc = face_center.get(key)
            if fc is None:
                fc = len(outV)
                outV.append(tuple(0.25 * (V[e[f[0]]] + V[e[f[1]]] + V[e[f[2]]] + V[e[f[3]]])))
                face_center[key] = fc
            for v in range(4):
                tets.append((int(e[f[(v + 1) % 4]]), int(e[f[v]]), fc, hc))
    return np.array(outV, dtype=float), np.array(tets, dtype=np.int64)


def quad_tri_subdiv(V, Q):
    """quad_tri_subdiv (filters/quad_tri_subdiv.hh): centre vertex, 4 triangles
    (e[v], e[v+1], centre)."""
    outV = [tuple(v) for v in V]
    tris = []
    for e in Q:
        c = len(outV)
        outV.append(tuple(np.sum(V[e], axis=0) / 4))
        for v in range(4):
            tris.append((int(e[v]), int(e[(v + 1) % 4]), c))
    return np.array(outV, dtype=float), np.array(tris, dtype=np.int64)


def grid_simplices(sizes, min_corner=None, max_corner=None):
    """`grid AxB[xC] -t [-m min -M max]` (src/bin/tools/grid.cc:115-137)."""
    V, E = gen_grid(sizes)
    if min_corner is not None:
        mn = np.zeros(3); mx = np.zeros(3)
        mn[:len(sizes)] = min_corner; mx[:len(sizes)] = max_corner
        scale = (mx - mn)
        for i in range(len(sizes)):
            scale[i] /= sizes[i]
        V = V * scale + mn
    if len(sizes) == 2:
        return quad_tri_subdiv(V, E)
    return hex_tet_subdiv(V, E)


# ----------------------------------------------------------------------------
# FEMMesh: node numbering + boundary extraction (FEMMesh.inl:11-82,
# TetMesh.inl:16-120, TriMesh.inl:16-140, FEMMesh.hh:221-237, 366-451)
# ----------------------------------------------------------------------------
HALF_FACE_CORNERS = ((1, 3, 2), (0, 2, 3), (0, 3, 1), (0, 1, 2))   # TetMesh.hh:221-226


@dataclass
class FEMMesh:
    N: int
    deg: int
    vertices: np.ndarray              # (nV, N)
    simplices: np.ndarray             # (ne, N+1) vertex indices
    nodes: np.ndarray = None          # (nNodes, N)
    elem_nodes: np.ndarray = None     # (ne, nodesPerElem) node indices, reference local order
    bdry_elem_vertices: np.ndarray = None   # (nbe, N) *volume vertex* indices, boundary corner order
    bdry_elem_nodes: np.ndarray = None      # (nbe, nodesPerBdryElem) *volume node* indices
    bdry_nodes: np.ndarray = None     # (nbn,) volume node index of each boundary node, reference order
    bdry_node_of_node: np.ndarray = None    # (nNodes,) boundary node index or -1
    vol: np.ndarray = None            # (ne,)
    G: np.ndarray = None              # (ne, N, N+1)
    bdry_vol: np.ndarray = None
    bdry_normal: np.ndarray = None
    bbox_min: np.ndarray = None
    bbox_max: np.ndarray = None

    @property
    def num_vertices(self):
        return self.vertices.shape[0]

    @property
    def num_nodes(self):
        return self.nodes.shape[0]

    @property
    def num_elements(self):
        return self.simplices.shape[0]


def build_mesh(N, deg, vertices, simplices):
    """FEMMesh constructor (FEMMesh.inl:11-82)."""
    V = np.asarray(vertices, dtype=float)[:, :N].copy()     # truncateFrom3D
    S = np.asarray(simplices, dtype=np.int64)
    assert S.shape[1] == N + 1
    if S.size and (S.min() < 0 or S.max() >= V.shape[0]):
        raise RuntimeError("Bad vertex index encountered.")
    m = FEMMesh(N=N, deg=deg, vertices=V, simplices=S)
    ne = S.shape[0]
    nV = V.shape[0]
    nedge = num_edges(N)

    # ---- volume edge nodes: first encounter over (element, local edge) (FEMMesh.inl:22-37)
    if deg == 2:
        a = S[:, list(EDGE_START[:nedge])].reshape(-1)
        b = S[:, list(EDGE_END[:nedge])].reshape(-1)
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        key = lo * nV + hi
        uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")       # unique ids sorted by first encounter
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        edge_id = rank[inv].reshape(ne, nedge)
        n_edge_nodes = uniq.size
        edge_lo = (uniq // nV)[order]
        edge_hi = (uniq % nV)[order]
        m.nodes = np.vstack([V, 0.5 * (V[edge_lo] + V[edge_hi])])     # FEMMesh.hh:228-233
        m.elem_nodes = np.hstack([S, nV + edge_id])
        edge_lookup = {(int(l), int(h)): i for i, (l, h) in enumerate(zip(edge_lo, edge_hi))}
    else:
        m.nodes = V.copy()
        m.elem_nodes = S.copy()

    # ---- boundary extraction
    if N == 3:
        # leftover half-faces of std::map<UnorderedTriplet,int>, iterated in sorted order
        # (TetMesh.inl:36-90)
        hf_v = np.stack([S[:, list(c)] for c in HALF_FACE_CORNERS], axis=1).reshape(-1, 3)  # hf = 4t+f
        srt = np.sort(hf_v, axis=1)
        key = (srt[:, 0] * nV + srt[:, 1]) * nV + srt[:, 2]
        uniq, idx, counts = np.unique(key, return_index=True, return_counts=True)
        if np.any(counts > 2):
            raise RuntimeError("Non-manifold input detected.")
        bhf = idx[counts == 1]               # np.unique sorts by key == lexicographic triple order
        # boundary face corner c = volume half-face corner 2 - c (TetMesh.hh:462-468)
        bverts = hf_v[bhf][:, ::-1]
        # boundary vertex numbering: first appearance over volume half-face corners 0..2
        # (TetMesh.inl:76-85 uses m_vertexOfHalfFace(c, bhf), c = 0,1,2)
        seq = hf_v[bhf].reshape(-1)
    else:
        # TriMesh.inl:38-117: half-edge he = 3t + c, TIP = corner (c+2)%3, TAIL = corner (c+1)%3
        tip = np.stack([S[:, (c + 2) % 3] for c in range(3)], axis=1).reshape(-1)
        tail = np.stack([S[:, (c + 1) % 3] for c in range(3)], axis=1).reshape(-1)
        lo, hi = np.minimum(tip, tail), np.maximum(tip, tail)
        key = lo * nV + hi
        uniq, idx, counts = np.unique(key, return_index=True, return_counts=True)
        if np.any(counts > 2):
            raise RuntimeError("Non-manifold edge detected")
        vhe = idx[counts == 1]
        # boundary edge tip = volume half-edge tail and vice versa (:104-106);
        # vertex(0) = tail(), vertex(1) = tip() (Handles/TriMeshHandles.hh:269)
        b_tip, b_tail = tail[vhe], tip[vhe]
        bverts = np.stack([b_tail, b_tip], axis=1)
        # boundary vertex creation order: tip then tail per boundary edge (:109-110)
        seq = np.stack([b_tip, b_tail], axis=1).reshape(-1)
    _, first = np.unique(seq, return_index=True)
    bV = seq[np.sort(first)]                 # boundary vertices in first-appearance order
    m.bdry_elem_vertices = bverts
    nbe = bverts.shape[0]

    if deg == 2:
        # boundary edge nodes: first encounter over boundary simplices x local edges
        # (FEMMesh.inl:39-59)
        nbedge = num_edges(N - 1)
        bedge_nodes = np.empty((nbe, nbedge), dtype=np.int64)
        seen = {}
        b_edge_list = []
        for be in range(nbe):
            for ei in range(nbedge):
                a_, b_ = int(bverts[be, EDGE_START[ei]]), int(bverts[be, EDGE_END[ei]])
                vn = edge_lookup[(min(a_, b_), max(a_, b_))]
                if vn not in seen:
                    seen[vn] = len(b_edge_list)
                    b_edge_list.append(vn)
                bedge_nodes[be, ei] = nV + vn
        m.bdry_elem_nodes = np.hstack([bverts, bedge_nodes]) if nbedge else bverts.copy()
        m.bdry_nodes = np.concatenate([bV, nV + np.array(b_edge_list, dtype=np.int64)])
    else:
        m.bdry_elem_nodes = bverts.copy()
        m.bdry_nodes = bV.copy()
    m.bdry_node_of_node = np.full(m.nodes.shape[0], -1, dtype=np.int64)
    m.bdry_node_of_node[m.bdry_nodes] = np.arange(m.bdry_nodes.size)

    # ---- embedding
    m.vol, m.G = embed_simplices(V[S])
    if nbe:
        m.bdry_vol, m.bdry_normal = embed_boundary(V[bverts])
    else:
        m.bdry_vol = np.zeros(0); m.bdry_normal = np.zeros((0, N))
    m.bbox_min = m.nodes.min(axis=0)
    m.bbox_max = m.nodes.max(axis=0)
    return m


# ----------------------------------------------------------------------------
# Per-element stiffness
# ----------------------------------------------------------------------------
def per_element_stiffness_reference_loops(N, deg, vol, G, D):
    """Literal restatement of Element::perElementStiffness for ONE element
    (LinearElasticity.hh:165-232), loop nest and all: returns the full matrix
    with ONLY the upper triangle written (strict lower triangle is NaN)."""
    nn = num_nodes(N, deg)
    C = tensor_C(N, D)
    T = grad_phi_interpolant(N, deg)                  # grad_phis[n] as interpolants
    gp = np.einsum("iva,ra->ivr", T, G)               # nodal values: (node, interp node, N)
    P, w = quadrature_points(N, 2 * (deg - 1))
    if deg == 1:
        ev = lambda f, p: f[0]
    else:
        ev = lambda f, p: np.tensordot(p, f, axes=(0, 0))   # linear interpolant at bary p
    Ke = np.full((N * nn, N * nn), np.nan)
    for c in range(N):
        for d in range(c, N):
            M = C[:, c, d, :]                          # M(a,b) = C(a,c,d,b)   :203-205
            for j in range(nn):
                vj = j * N + d
                Mgpj = gp[j] @ M.T                     # :212-213
                for i in range(nn):
                    vi = i * N + c
                    if c == d and vi > vj:
                        continue
                    val = 0.0
                    for q in range(len(w)):
                        val += w[q] * float(np.dot(ev(gp[i], P[q]), ev(Mgpj, P[q])))
                    val *= vol
                    if vi <= vj:
                        Ke[vi, vj] = val
                    else:
                        Ke[vj, vi] = val
    return Ke


def per_element_stiffness(N, deg, vol, G, D):
    """Vectorised Ke = sum_q w_q vol B_q^T D B_q in the (i,c),(j,d) layout of
    LinearElasticity.hh:165-232 (full symmetric matrices).  vol (ne,), G (ne,N,N+1),
    D (flat,flat) or (ne,flat,flat).  Returns (ne, N*nn, N*nn)."""
    ne = vol.shape[0]
    nn = num_nodes(N, deg)
    D = np.asarray(D, dtype=float)
    P, w = quadrature_points(N, 2 * (deg - 1))
    Ke = np.zeros((ne, nn, N, nn, N))
    if D.ndim == 2:
        C = tensor_C(N, D)
        for q in range(len(w)):
            A = grad_phi_coeffs(N, deg, P[q])            # (nn, N+1)
            g = np.einsum("ia,era->eir", A, G)           # grad phi_i at x_q: (ne, nn, N)
            # Ke[(i,c),(j,d)] += w vol g_i[a] C(a,c,d,b) g_j[b]
            H = np.einsum("acdb,ejb->ejacd", C, g)
            Ke += w[q] * np.einsum("eia,ejacd->eicjd", g, H)
    else:
        for q in range(len(w)):
            A = grad_phi_coeffs(N, deg, P[q])
            g = np.einsum("ia,era->eir", A, G)
            Cq = np.zeros((ne, N, N, N, N))
            for a in range(N):
                for c in range(N):
                    for d in range(N):
                        for b in range(N):
                            Cq[:, a, c, d, b] = D[:, flatten_indices(N, a, c), flatten_indices(N, d, b)]
            H = np.einsum("eacdb,ejb->ejacd", Cq, g)
            Ke += w[q] * np.einsum("eia,ejacd->eicjd", g, H)
    Ke *= vol[:, None, None, None, None]
    return Ke.reshape(ne, nn * N, nn * N)


def per_element_stiffness_BtDB(N, deg, vol, G, D):
    """Third, independent formulation: explicit engineering-strain B matrix
    (SURVEY B.1).  One element.  Used to cross-check the other two."""
    nn = num_nodes(N, deg)
    nf = flat_len(N)
    P, w = quadrature_points(N, 2 * (deg - 1))
    Ke = np.zeros((N * nn, N * nn))
    for q in range(len(w)):
        A = grad_phi_coeffs(N, deg, P[q])
        g = A @ G.T                                         # (nn, N)
        B = np.zeros((nf, N * nn))
        for i in range(nn):
            for c in range(N):
                for var in range(N):
                    f = flatten_indices(N, c, var)
                    B[f, N * i + c] += g[i, var] if var != c else 0.0
                B[c, N * i + c] = g[i, c]
        Ke += w[q] * vol * (B.T @ D @ B)
    return Ke


# ----------------------------------------------------------------------------
# Global assembly: upper-triangle triplets in DoF space
# (LinearElasticity.hh:1408-1466) + sumRepeated/CSC (SparseMatrices.hh:280-374)
# ----------------------------------------------------------------------------
def assemble_upper_triplets(mesh, D, dof_for_node=None, Ke=None):
    N, deg = mesh.N, mesh.deg
    nn = mesh.elem_nodes.shape[1]
    if Ke is None:
        Ke = per_element_stiffness(N, deg, mesh.vol, mesh.G, D)
    dof = mesh.elem_nodes if dof_for_node is None else np.asarray(dof_for_node)[mesh.elem_nodes]
    var = (N * dof[:, :, None] + np.arange(N)[None, None, :]).reshape(mesh.num_elements, nn * N)
    I = np.broadcast_to(var[:, :, None], Ke.shape)
    J = np.broadcast_to(var[:, None, :], Ke.shape)
    di = np.broadcast_to(np.repeat(dof, N, axis=1)[:, :, None], Ke.shape)
    dj = np.broadcast_to(np.repeat(dof, N, axis=1)[:, None, :], Ke.shape)
    keep = (di <= dj) & (I <= J)                       # :1421, :1425
    return I[keep], J[keep], Ke[keep]


def upper_csc(n, I, J, V):
    """TripletMatrix::sumRepeated + CSC (SparseMatrices.hh:280-374, 423-447)."""
    A = sp.coo_matrix((V, (I, J)), shape=(n, n)).tocsc()
    A.sum_duplicates()
    A.sort_indices()
    return A


def full_symmetric(Aupper):
    return (Aupper + sp.triu(Aupper, 1).T).tocsr()


def stiffness_matrix(mesh, D, dof_for_node=None, num_dofs=None):
    """Full symmetric K (CSR) in DoF space."""
    nd = mesh.num_nodes if dof_for_node is None else num_dofs
    I, J, V = assemble_upper_triplets(mesh, D, dof_for_node)
    return full_symmetric(upper_csc(mesh.N * nd, I, J, V))


# ----------------------------------------------------------------------------
# SPSDSystem semantics (SparseMatrices.hh:2389-2500, 2516-2606)
# ----------------------------------------------------------------------------
def solve_fixed(K, f, fixed_vars, fixed_vals):
    """K_ff u_f = f_f - K_fc u_c ; u_c = g.  Direct (SuperLU stands in for CHOLMOD)."""
    n = K.shape[0]
    fixed_vars = np.asarray(fixed_vars, dtype=np.int64)
    if len(set(fixed_vars.tolist())) != fixed_vars.size:
        raise RuntimeError("Variable already fixed.")
    if f.shape[0] != n:
        raise RuntimeError("Bad RHS")
    free = np.ones(n, dtype=bool)
    free[fixed_vars] = False
    u = np.zeros(n)
    u[fixed_vars] = fixed_vals
    Kcsr = K.tocsr()
    b = f[free] - (Kcsr[free][:, ~free] @ u[~free])
    Kff = Kcsr[free][:, free].tocsc()
    # symmetric-mode minimum-degree ordering: ~40x less fill than SuperLU's default COLAMD on 3D elasticity
    lu = spla.splu(Kff, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    u[free] = lu.solve(b)
    return u


def solve_constrained(K, C, C_rhs, f, fixed_vars, fixed_vals):
    """SPSDSystem::setConstrained + fixVariables + solve with Lagrange rows (SparseMatrices.hh:2332-2348,
    2389-2500, 2516-2606): the saddle-point matrix [[K, C^T], [C, 0]] with the fixed variables' rows and
    columns removed (their columns moved to the right-hand side, also for the constraint rows), solved
    directly (sparse LU stands in for UMFPACK).  Returns (u, multipliers)."""
    n = K.shape[0]
    C = sp.csr_matrix(np.atleast_2d(np.asarray(C, dtype=float)))
    m = C.shape[0]
    if C.shape[1] != n or len(C_rhs) != m:
        raise RuntimeError("Bad constraint rows")
    if np.asarray(f).reshape(-1).shape[0] != n:
        raise RuntimeError("Bad RHS")
    fixed_vars = np.asarray(fixed_vars, dtype=np.int64)
    if len(set(fixed_vars.tolist())) != fixed_vars.size:
        raise RuntimeError("Variable already fixed.")
    free = np.ones(n, dtype=bool)
    free[fixed_vars] = False
    u = np.zeros(n)
    u[fixed_vars] = fixed_vals
    Kcsr = K.tocsr()
    b = np.concatenate([np.asarray(f, float).reshape(-1)[free] - Kcsr[free][:, ~free] @ u[~free],
                        np.asarray(C_rhs, float) - C[:, ~free] @ u[~free]])
    Cf = C[:, free]
    A = sp.bmat([[Kcsr[free][:, free], Cf.T], [Cf, None]], format="csc")
    x = spla.splu(A).solve(b)
    u[free] = x[:free.sum()]
    return u, x[free.sum():]


# ----------------------------------------------------------------------------
# Loads and post-processing
# ----------------------------------------------------------------------------
def neumann_load(mesh, tractions, dof_for_node=None, num_dofs=None, delta_forces=()):
    """Simulator::neumannLoad (LinearElasticity.hh:703-717) with
    BoundaryElement::nodalNeumannLoad (:341-347)."""
    nd = mesh.num_nodes if dof_for_node is None else num_dofs
    dof = (lambda n: n) if dof_for_node is None else (lambda n: dof_for_node[n])
    load = np.zeros((nd, mesh.N))
    wts = integrated_phis(mesh.N - 1, mesh.deg)
    for be in range(mesh.bdry_elem_nodes.shape[0]):
        for n in range(mesh.bdry_elem_nodes.shape[1]):
            load[dof(mesh.bdry_elem_nodes[be, n])] += wts[n] * mesh.bdry_vol[be] * tractions[be]
    for node, fvec in delta_forces:
        load[dof(node)] += fvec
    return load


def sym_to_flat(N, M):
    out = np.zeros(flat_len(N))
    for i in range(flat_len(N)):
        a, b = unflatten_index(N, i)
        out[i] = M[a, b]
    return out


def flat_to_sym(N, f):
    M = np.zeros((N, N))
    for i in range(flat_len(N)):
        a, b = unflatten_index(N, i)
        M[a, b] = M[b, a] = f[i]
    return M


def constant_strain_load(mesh, D, strain_flat, dof_for_node=None, num_dofs=None):
    """Simulator::constantStrainLoad (LinearElasticity.hh:551-562) with
    perElementConstantStrainLoad/StressLoad (:135-162): l_i = (C:eps) int grad phi_i."""
    N, deg = mesh.N, mesh.deg
    nd = mesh.num_nodes if dof_for_node is None else num_dofs
    D = np.asarray(D)
    if D.ndim == 2:
        sig = np.broadcast_to(flat_to_sym(N, double_contract(N, D, np.asarray(strain_flat))),
                              (mesh.num_elements, N, N))
    else:
        sflat = np.stack([double_contract(N, D[e], np.asarray(strain_flat)) for e in range(D.shape[0])])
        sig = np.stack([flat_to_sym(N, s) for s in sflat])
    T = grad_phi_interpolant(N, deg)          # (nn, nv, N+1)
    Tint = T.mean(axis=1)                     # integrate interpolant / vol
    gint = np.einsum("ia,era->eir", Tint, mesh.G) * mesh.vol[:, None, None]   # (ne, nn, N)
    l = np.einsum("ecr,eir->eic", sig, gint)
    dof = mesh.elem_nodes if dof_for_node is None else np.asarray(dof_for_node)[mesh.elem_nodes]
    load = np.zeros((nd, N))
    np.add.at(load, dof.reshape(-1), l.reshape(-1, N))
    return load


def element_strain_vertices(mesh, u_nodes):
    """Element::strain (LinearElasticity.hh:99-116): strain interpolant nodal values,
    returns (ne, nInterp, N, N) with nInterp = 1 (deg 1) or N+1 (deg 2)."""
    N, deg = mesh.N, mesh.deg
    T = grad_phi_interpolant(N, deg)
    gp = np.einsum("iva,era->eivr", T, mesh.G)             # (ne, nn, nv, N)
    ue = u_nodes[mesh.elem_nodes]                          # (ne, nn, N)
    grad = np.einsum("eic,eivr->evcr", ue, gp)             # du_c/dx_r
    return 0.5 * (grad + np.swapaxes(grad, 2, 3))


def average_strain_stress(mesh, D, u_nodes):
    """averageStrainField / averageStressField (LinearElasticity.hh:528-549):
    mean of the interpolant's nodal values; flattened Voigt (ne, flat)."""
    N = mesh.N
    eps = element_strain_vertices(mesh, u_nodes).mean(axis=1)
    flat = np.stack([eps[:, a, b] for a, b in (unflatten_index(N, i) for i in range(flat_len(N)))], axis=1)
    D = np.asarray(D)
    if D.ndim == 2:
        stress = double_contract(N, D, flat)
    else:
        e2 = flat.copy(); e2[:, N:] *= 2
        stress = np.einsum("eij,ej->ei", D, e2)
    return flat, stress


def apply_stiffness_matrix(mesh, D, u_nodes):
    """Simulator::applyStiffnessMatrix (LinearElasticity.hh:801-823): raw K on nodes
    (ignores periodic DoFs and Dirichlet)."""
    N = mesh.N
    Ke = per_element_stiffness(N, mesh.deg, mesh.vol, mesh.G, D)
    ue = u_nodes[mesh.elem_nodes].reshape(mesh.num_elements, -1)
    fe = np.einsum("eij,ej->ei", Ke, ue).reshape(mesh.num_elements, -1, N)
    out = np.zeros_like(u_nodes)
    np.add.at(out, mesh.elem_nodes.reshape(-1), fe.reshape(-1, N))
    return out


# ----------------------------------------------------------------------------
# Boundary conditions: .bc parsing (BoundaryConditions.cc:217-388) and their
# application (LinearElasticity.hh:881-1111, 1169-1249, 1469-1518, 1595-1618)
# ----------------------------------------------------------------------------
def _parse_vector_lenient(v):
    """parseVectorLenient (BoundaryConditions.cc:28-45): 2- or 3-vectors of numbers."""
    if not isinstance(v, list) or len(v) not in (2, 3):
        raise ValueError("Error parsing vector; read %d components" % (len(v) if isinstance(v, list) else -1))
    out = np.zeros(3)
    for i, x in enumerate(v):
        if isinstance(x, bool) or not isinstance(x, (int, float)):
            raise ValueError("Error parsing vector")
        out[i] = x
    return out


_EXPR_FUNCS = {k: getattr(math, k) for k in
               ("acos", "asin", "atan", "atan2", "ceil", "cos", "cosh", "exp", "floor", "log10",
                "pow", "sin", "sinh", "sqrt", "tan", "tanh")}
_EXPR_FUNCS.update({"abs": abs, "ln": math.log, "log": math.log10, "fac": lambda a: math.gamma(a + 1),
                    "ncr": lambda n, r: math.comb(int(n), int(r)), "npr": lambda n, r: math.perm(int(n), int(r)),
                    "pi": math.pi, "e": math.e})


def eval_expression(expr, env):
    """tinyexpr-compatible evaluation (ExpressionVector.hh:21-58); `^` is power."""
    ns = dict(_EXPR_FUNCS)
    ns.update(env)
    return float(eval(expr.replace("^", "**"), {"__builtins__": {}}, ns))


@dataclass
class Condition:
    kind: str                      # dirichlet | target | traction | force | pressure | delta force | ... nodes/elements
    region_min: np.ndarray = None
    region_max: np.ndarray = None
    value: np.ndarray = None       # plain value (N,)
    exprs: list = None             # expression strings (len N) when value is an expression vector
    cmask: tuple = (True, True, True)
    node_indices: list = None
    node_values: list = None
    elem_corners: list = None
    elem_values: list = None
    elem_vertices: list = None


def _component_mask(s):
    return tuple(ch in s for ch in "xyz")


def read_boundary_conditions(N, params, bbox_min, bbox_max):
    """readBoundaryConditions<N> (BoundaryConditions.cc:217-388).  `params` is the parsed
    JSON.  Returns (conds, noRigidMotion, periodicPairs, pinTranslationMask)."""
    no_rigid = bool(params.get("no_rigid_motion", False))
    pps = []
    comp = ["x", "y", "z"]
    for c in range(N):
        key = "fix_periodic_pair_" + comp[c]
        if key in params:
            spec = params[key]
            face = N
            for c2 in range(N):
                if c2 != c and spec == comp[c2]:
                    face = c2
            if face == N:
                raise RuntimeError("invalid " + key)
            pps.append((c, face))
    pin = _component_mask(params.get("pin_translation", ""))
    conds = []
    bmin = np.asarray(bbox_min, float); bmax = np.asarray(bbox_max, float)
    for t in params["regions"]:
        typ = t["type"]
        cmask = (True, True, True)
        prefix = ""
        if typ[:9] == "dirichlet":
            prefix, typ = "dirichlet", typ[9:]
        elif typ[:6] == "target":
            prefix, typ = "target", typ[6:]
        if prefix:
            ln = 0
            for ch in typ:
                if ch < "x" or ch > "z":
                    break
                ln += 1
            if ln > 3:
                raise RuntimeError("invalid mask")
            if ln > 0:
                cmask = _component_mask(typ[:ln])
            typ = prefix + typ[ln:]
        c = Condition(kind=typ, cmask=cmask)
        if "nodes" in typ:
            c.node_indices, c.node_values = [], []
            for val in t["values"]:
                disp = _parse_vector_lenient(val[0])[:N]
                for nd in val[1]:
                    c.node_indices.append(int(nd)); c.node_values.append(disp.copy())
        elif typ in ("traction elements", "pressure elements", "force elements"):
            c.elem_corners, c.elem_values = [], []
            for val in t["values"]:
                vec = _parse_vector_lenient(val[0])[:N]
                for elem in val[1]:
                    idx = [int(x) for x in elem]
                    if len(idx) == 2:
                        idx.append(0)
                    if len(idx) != 3:
                        raise RuntimeError("Error parsing element condition values.")
                    c.elem_corners.append(tuple(sorted(idx))); c.elem_values.append(vec.copy())
        else:
            c.region_min = np.zeros(N); c.region_max = np.zeros(N)
            if "box" in t:
                c.region_min = _parse_vector_lenient(t["box"]["minCorner"])[:N]
                c.region_max = _parse_vector_lenient(t["box"]["maxCorner"])[:N]
            elif "box%" in t:
                rmin = _parse_vector_lenient(t["box%"]["minCorner"])[:N]
                rmax = _parse_vector_lenient(t["box%"]["maxCorner"])[:N]
                c.region_min = bmin + rmin * (bmax - bmin)          # Geometry.hh:259-262
                c.region_max = bmin + rmax * (bmax - bmin)
            elif "element vertices" in t:
                c.elem_vertices = [tuple(int(x) for x in ev) for ev in t["element vertices"]]
            elif "path" in t or "polygon" in t:
                raise NotImplementedError("path/polygon regions are outside the oracle's scope")
            try:
                c.value = _parse_vector_lenient(t["value"])[:N]
            except ValueError:
                ex = []
                for v in t["value"]:
                    if isinstance(v, str):
                        ex.append(v)
                    elif isinstance(v, (int, float)):
                        ex.append(repr(v))
                    else:
                        raise RuntimeError("Failed to parse expression vector")
                if N == 2 and len(ex) == 3 and float(ex[2]) == 0:
                    ex.pop()
                if len(ex) != N:
                    raise RuntimeError("Incorrect expression vector size")
                c.exprs = ex
                if typ not in ("traction", "dirichlet", "dirichlet elements", "target", "delta force"):
                    raise RuntimeError("Only region-based traction, dirichlet, target, and delta force support "
                                       "expression vectors")
        known = ("pressure", "traction", "force", "dirichlet", "dirichlet elements", "target", "contact",
                 "fracture", "dirichlet nodes", "target nodes", "traction elements", "pressure elements",
                 "force elements", "delta force", "delta force nodes")
        if typ not in known:
            raise RuntimeError("Invalid type '%s'" % typ)
        conds.append(c)
    return conds, no_rigid, pps, pin


class Simulator:
    """CPU oracle mirror of LinearElasticity::Simulator (LinearElasticity.hh:434-1659)
    restricted to the assemble-and-solve path."""

    def __init__(self, N, deg, vertices, simplices):
        self.mesh = build_mesh(N, deg, vertices, simplices)
        neg = int((self.mesh.vol < 0).sum())
        if neg > 0:                                       # :465-472
            raise RuntimeError("Mesh has negatively oriented elements.\n"
                               "Correct with: mesh_convert --reorientNegativeElements.")
        self.N, self.deg = N, deg
        m = self.mesh
        nbn = m.bdry_nodes.size
        self.D = isotropic_D(N, 1.0, 0.3)                 # Materials.hh:408 default
        self.dirichlet_comp = np.zeros((nbn, N), dtype=bool)
        self.dirichlet_disp = np.zeros((nbn, N))
        self.neumann_traction = np.zeros((m.bdry_elem_nodes.shape[0], N))
        self.delta_forces = []
        self.dof_for_node = None
        self.num_dofs_ = m.num_nodes
        self.use_rigid_motion_constraint = False
        self.use_nrt_pin = False
        self.is_internal_be = np.zeros(m.bdry_elem_nodes.shape[0], dtype=bool)

    # -- DoF bookkeeping (:825-836)
    def num_dofs(self):
        return self.num_dofs_ if self.dof_for_node is not None else self.mesh.num_nodes

    def DoF(self, node):
        return int(self.dof_for_node[node]) if self.dof_for_node is not None else int(node)

    def set_material(self, D):
        self.D = np.asarray(D, dtype=float)

    # -- BoundaryNode::setDirichlet (:392-406)
    def _set_dirichlet(self, bn, mask, val):
        for c in range(self.N):
            if not mask[c]:
                continue
            if not self.dirichlet_comp[bn, c]:
                self.dirichlet_comp[bn, c] = True
                self.dirichlet_disp[bn, c] = val[c]
            elif abs(self.dirichlet_disp[bn, c] - val[c]) > 1e-10:
                raise RuntimeError("Conflicting dirichlet displacements.")

    def apply_translation_pins(self, mask):               # :1095-1111
        m = self.mesh
        for d in range(self.N):
            if not mask[d]:
                continue
            p = m.nodes[m.bdry_nodes, d]
            bn = int(np.argmin(p))                        # first minimum in boundary-node order
            cm = [False] * 3; cm[d] = True
            self._set_dirichlet(bn, cm, np.zeros(self.N))

    def apply_boundary_conditions(self, conds):           # :881-1027
        m = self.mesh; N = self.N
        env0 = {}
        dims = m.bbox_max - m.bbox_min
        for i in range(N):
            env0["mesh_size_%d" % i] = dims[i]; env0["mesh_min_%d" % i] = m.bbox_min[i]
            env0["mesh_max_%d" % i] = m.bbox_max[i]

        def env_for(cond, p):
            env = dict(env0)
            if cond.region_min is not None:
                for i in range(N):
                    env["region_size_%d" % i] = cond.region_max[i] - cond.region_min[i]
                    env["region_min_%d" % i] = cond.region_min[i]; env["region_max_%d" % i] = cond.region_max[i]
            env["x"] = p[0]; env["y"] = p[1]; env["z"] = p[2] if N == 3 else 0.0
            return env

        def value_at(cond, p):
            if cond.exprs is None:
                return cond.value
            env = env_for(cond, p)
            return np.array([eval_expression(e, env) for e in cond.exprs])

        def contains(cond, p):                            # Geometry.hh:276-279 inclusive
            return bool(np.all(p >= cond.region_min) and np.all(p <= cond.region_max))

        for cond in conds:
            k = cond.kind
            if k in ("traction", "force", "pressure"):
                area = 0.0; region = []
                for be in range(m.bdry_elem_vertices.shape[0]):
                    center = m.nodes[m.bdry_elem_vertices[be]].mean(axis=0)     # :903-906
                    if contains(cond, center):
                        area += m.bdry_vol[be]; region.append(be)
                        if k == "pressure":
                            self.neumann_traction[be] = -cond.value[0] * m.bdry_normal[be]
                        else:
                            self.neumann_traction[be] = value_at(cond, center)
                if not region:
                    raise RuntimeError("Neumann region unmatched")
                if k == "force":
                    self.neumann_traction[region] /= area
            elif k in ("target", "target nodes"):
                pass                                      # warned and ignored, :934-939
            elif k == "dirichlet":
                for bn, node in enumerate(m.bdry_nodes):
                    p = m.nodes[node]
                    if contains(cond, p):
                        self._set_dirichlet(bn, cond.cmask, value_at(cond, p))
            elif k == "dirichlet elements":
                evs = set(cond.elem_vertices)
                for be in range(m.bdry_elem_vertices.shape[0]):
                    if tuple(int(v) for v in m.bdry_elem_vertices[be]) in evs:
                        for node in m.bdry_elem_nodes[be]:
                            self._set_dirichlet(int(m.bdry_node_of_node[node]), cond.cmask,
                                                value_at(cond, m.nodes[node]))
            elif k in ("traction elements", "pressure elements", "force elements"):
                table = {c: v for c, v in zip(cond.elem_corners, cond.elem_values)}
                area = 0.0; region = []; nset = 0
                for be in range(m.bdry_elem_vertices.shape[0]):
                    bv = [int(v) for v in m.bdry_elem_vertices[be]]
                    key = tuple(sorted(bv + ([0] if N == 2 else [])))
                    if key in table:
                        val = table[key]
                        if k == "pressure elements":
                            self.neumann_traction[be] = -val[0] * m.bdry_normal[be]
                        else:
                            self.neumann_traction[be] = val
                            if k == "force elements":
                                area += m.bdry_vol[be]; region.append(be)
                        nset += 1
                if nset != len(table):
                    raise RuntimeError("Some element boundary conditions weren't matched.")
                if region:
                    self.neumann_traction[region] /= area
            elif k == "dirichlet nodes":
                for ni, val in zip(cond.node_indices, cond.node_values):
                    bn = int(m.bdry_node_of_node[ni])
                    if bn < 0:
                        raise RuntimeError("Condition applied to non-boundary node %d" % ni)
                    self._set_dirichlet(bn, cond.cmask, val)
            elif k == "delta force":
                for node in range(m.num_nodes):
                    if contains(cond, m.nodes[node]):
                        self.delta_forces.append((node, value_at(cond, m.nodes[node])))
            elif k == "delta force nodes":
                for ni, val in zip(cond.node_indices, cond.node_values):
                    if ni > m.num_nodes:                 # sic, :1021
                        raise RuntimeError("DeltaForceNodesCondition node index out of bounds: %d" % ni)
                    self.delta_forces.append((ni, val))
            else:
                raise RuntimeError("Illegal BC type")

    def apply_no_rigid_motion_constraint(self):           # :1052-1059
        self.use_rigid_motion_constraint = True

    def set_use_pin_no_rigid_translation_constraint(self, use):   # :1063
        self.use_nrt_pin = use

    def set_periodic(self, dof_for_node, num_dofs, is_periodic_be):    # applyPeriodicConditions :845-854
        self.dof_for_node = np.asarray(dof_for_node, dtype=np.int64)
        self.num_dofs_ = int(num_dofs)
        self.is_internal_be = np.asarray(is_periodic_be, dtype=bool)

    # -- constraint bookkeeping
    def _pin_node(self, comps=(True, True, True)):        # m_pinNode :1595-1618
        m = self.mesh
        interior = np.nonzero(m.bdry_node_of_node < 0)[0]
        node = int(interior[0]) if interior.size else 0
        vars_ = [self.N * self.DoF(node) + d for d in range(self.N) if comps[d]]
        return vars_, [0.0] * len(vars_)

    def _dirichlet_vars_and_values(self):                 # m_getDirichletVarsAndValues :1469-1518
        m = self.mesh; N = self.N
        cidx = {}
        cdofs, cdisp, ccomp = [], [], []
        for bn, node in enumerate(m.bdry_nodes):
            if not self.dirichlet_comp[bn].any():
                continue
            dof = self.DoF(node)
            if dof not in cidx:
                cidx[dof] = len(cdofs)
                cdofs.append(dof); cdisp.append(self.dirichlet_disp[bn].copy()); ccomp.append(self.dirichlet_comp[bn].copy())
            else:
                k = cidx[dof]
                if np.linalg.norm(self.dirichlet_disp[bn] - cdisp[k]) > 1e-10 or \
                        not np.array_equal(self.dirichlet_comp[bn], ccomp[k]):
                    raise RuntimeError("Mismatched Dirichlet constraint on periodic DoF")
        vars_, vals = [], []
        for dof, disp, comp in zip(cdofs, cdisp, ccomp):
            for c in range(N):
                if comp[c]:
                    vars_.append(N * dof + c); vals.append(disp[c])
        return vars_, vals

    # -- Lagrange rows (m_appendInfinitesimalRotationMatrix :1530-1568, m_appendTranslationMatrix :1571-1593)
    def _rotation_rows(self):
        N, m = self.N, self.mesh
        nd, nn = self.num_dofs(), m.num_nodes
        if N == 2 and nd < nn:                            # periodic conditions pin the rotations (:1539)
            return np.zeros((0, N * nd))
        if nd < nn - 1:
            return np.zeros((0, N * nd))
        if nd < nn:
            raise RuntimeError("Single pair periodic BC unsupported in 3D.")
        X = m.nodes
        if N == 3:
            R = np.zeros((3, nn, 3))
            R[0, :, 1], R[0, :, 2] = -X[:, 2], X[:, 1]    # (0, -z, y)
            R[1, :, 0], R[1, :, 2] = X[:, 2], -X[:, 0]    # (z, 0, -x)
            R[2, :, 0], R[2, :, 1] = -X[:, 1], X[:, 0]    # (-y, x, 0)
            return R.reshape(3, -1)
        R = np.zeros((1, nn, 2))
        R[0, :, 0], R[0, :, 1] = -X[:, 1], X[:, 0]
        return R.reshape(1, -1)

    def _translation_rows(self, comps=(True, True, True)):
        N, nd = self.N, self.num_dofs()
        rows = []
        for c in range(N):
            if comps[c]:
                T = np.zeros((nd, N)); T[:, c] = 1.0
                rows.append(T.reshape(-1))
        return np.array(rows).reshape(len(rows), N * nd)

    def rigid_mode_matrix(self):                          # m_assembleRigidModeMatrix :1522-1528 (rotations first)
        return np.vstack([self._rotation_rows(), self._translation_rows()])

    def rigid_inner_product(self, u_dofs):                # getRigidInnerProduct :1114-1126
        return self.rigid_mode_matrix() @ np.asarray(u_dofs, float).reshape(-1)

    def apply_rigid_motion_constraint(self, u_dofs):      # :1069-1076
        self.apply_no_rigid_motion_constraint()
        self.rigid_motion_rhs = self.rigid_inner_product(u_dofs)

    def constraints(self, allow_ill_posed=False):
        """assembleConstrainedSystem's constraint half (:1201-1249): fixed variables and values plus the
        Lagrange-multiplier rows C and their right-hand side."""
        N = self.N
        fixed, vals = [], []
        C = np.zeros((0, N * self.num_dofs()))
        rhs = np.zeros(0)
        if self.use_rigid_motion_constraint:
            C = self._rotation_rows()
            if self.use_nrt_pin:
                v, x = self._pin_node()
                fixed += v; vals += x
            else:
                C = np.vstack([C, self._translation_rows()])
            rhs = np.asarray(getattr(self, "rigid_motion_rhs", np.zeros(0)), float)
            if rhs.size == 0:
                rhs = np.zeros(C.shape[0])
            if rhs.size != C.shape[0]:
                raise RuntimeError("Invalid rigid motion RHS")
        elif not allow_ill_posed:
            needs_t = [not self.dirichlet_comp[:, c].any() for c in range(N)]      # :1169-1190
            if any(needs_t):
                if self.use_nrt_pin:
                    v, x = self._pin_node(tuple(needs_t) + (False,) * (3 - N))
                    fixed += v; vals += x
                else:
                    C = self._translation_rows(tuple(needs_t) + (False,) * (3 - N))
                    rhs = np.zeros(C.shape[0])
            if self.dirichlet_comp.sum() == 0:
                raise RuntimeError("Unimplemented")                                  # :1240
        v, x = self._dirichlet_vars_and_values()
        fixed += v; vals += x
        return np.array(fixed, dtype=np.int64), np.array(vals, dtype=float), C, rhs

    def fixed_vars_and_values(self, allow_ill_posed=False):
        fixed, vals, _, _ = self.constraints(allow_ill_posed)
        return fixed, vals

    # -- loads / solve
    def neumann_load(self):                               # :703-717
        return neumann_load(self.mesh, self.neumann_traction, self.dof_for_node, self.num_dofs(),
                            self.delta_forces)

    def constant_strain_load(self, strain_flat):          # :551-562
        return constant_strain_load(self.mesh, self.D, strain_flat, self.dof_for_node, self.num_dofs())

    def stiffness(self):
        return stiffness_matrix(self.mesh, self.D, self.dof_for_node, self.num_dofs())

    def dof_to_node_field(self, x):                       # :665-677
        x = np.asarray(x).reshape(-1, self.N)
        if self.dof_for_node is None:
            return x[:self.mesh.num_nodes].copy()
        return x[self.dof_for_node]

    def solve(self, f=None):                              # :479-487, 657
        if f is None:
            f = self.neumann_load()
        if not hasattr(self, "_K"):
            self._K = self.stiffness()
        fixed, vals, C, rhs = self.constraints()
        if C.shape[0]:
            x, self.multipliers = solve_constrained(self._K, C, rhs, np.asarray(f, float).reshape(-1), fixed, vals)
        else:
            x = solve_fixed(self._K, np.asarray(f, float).reshape(-1), fixed, vals)
        return self.dof_to_node_field(x)

    def average_strain_stress(self, u):
        return average_strain_stress(self.mesh, self.D, u)

    def apply_stiffness_matrix(self, u):
        return apply_stiffness_matrix(self.mesh, self.D, u)


def simulate(N, deg, vertices, simplices, D, bc_params):
    """What Simulate_cli does between reading its inputs and writing its fields
    (src/bin/Simulate_cli.cc:86-242)."""
    sim = Simulator(N, deg, vertices, simplices)
    sim.set_material(D)
    conds, no_rigid, pps, pin = read_boundary_conditions(N, bc_params, sim.mesh.bbox_min, sim.mesh.bbox_max)
    if pps:
        raise NotImplementedError("fix_periodic_pair_* is outside the oracle's scope")
    sim.apply_translation_pins(pin)
    sim.apply_boundary_conditions(conds)
    if no_rigid:
        sim.apply_no_rigid_motion_constraint()
    u = sim.solve()
    strain, stress = sim.average_strain_stress(u)
    load = sim.dof_to_node_field(sim.neumann_load())
    return dict(sim=sim, u=u, strain=strain, stress=stress, load=load, Ku=sim.apply_stiffness_matrix(u))


# ----------------------------------------------------------------------------
# Periodic conditions (PeriodicBoundaryMatcher.hh:38-258, BoundaryConditions.hh:457-561)
# and periodic homogenization (PeriodicHomogenization.hh:34-186)
# ----------------------------------------------------------------------------
def _match_periodic(P, lo, hi, on_min, on_max, eps):
    """PeriodicBoundaryMatcher::match (:149-258): one node set per minimal node, 2^(#faces) identified copies."""
    N = P.shape[1]
    minimal = ~on_max.any(axis=1)

    def find(q):
        d = np.linalg.norm(P - q, axis=1)
        d[minimal] = np.inf
        j = int(np.argmin(d))
        return j if d[j] <= eps else -1
    node_set_for = np.full(P.shape[0], -1, dtype=np.int64)
    node_sets = []
    for i in range(P.shape[0]):
        if not minimal[i]:
            continue
        node_set_for[i] = len(node_sets)
        faces = [d for d in range(N) if on_min[i, d]]
        ns = [i]
        for n in range(1, 1 << len(faces)):
            q = P[i].copy()
            for idx, d in enumerate(faces):
                if n & (1 << idx):
                    q[d] = hi[d]
            j = find(q)
            if j < 0:
                raise RuntimeError("Couldn't find %dth periodic-identified node for minimal boundary node %d" % (n, i))
            if node_set_for[j] != -1:
                raise RuntimeError("Non bijective node set assignment.")
            node_set_for[j] = node_set_for[i]
            ns.append(j)
        node_sets.append(ns)
    if (node_set_for < 0).any():
        raise RuntimeError("Unmatched non-minimal boundary node")
    return node_sets, node_set_for


def _match_periodic_permitting_mismatch(P, lo, hi, on_min, on_max, eps):
    """PeriodicBoundaryMatcher::matchPermittingMismatch (:268-372): per-axis pairing of max-face nodes with the
    min-face node at the projected position (if any); node sets = connected components in boundary-node order.
    Returns (node_sets, node_set_for, num_mismatches)."""
    nb, N = P.shape
    pair = np.full((nb, N), -1, dtype=np.int64)
    for d in range(N):
        cand = np.nonzero(on_min[:, d])[0]
        for i in np.nonzero(on_max[:, d])[0]:
            if cand.size == 0:
                continue
            q = P[i].copy(); q[d] = lo[d]
            dist = np.linalg.norm(P[cand] - q, axis=1)
            j = int(np.argmin(dist))
            if dist[j] > eps:
                continue
            pi = int(cand[j])
            if pair[i, d] != -1 or pair[pi, d] != -1:
                raise RuntimeError("Non-bijective boundary matching")
            pair[i, d] = pi; pair[pi, d] = i
    node_set_for = np.full(nb, -1, dtype=np.int64)
    node_sets = []
    mismatches = 0
    for i in range(nb):
        if node_set_for[i] != -1:
            continue
        nsi = len(node_sets)
        node_set_for[i] = nsi
        ns = [i]
        head = 0
        while head < len(ns):
            u = ns[head]; head += 1
            for d in range(N):
                if not (on_min[u, d] or on_max[u, d]):
                    continue
                v = pair[u, d]
                if v == -1:
                    mismatches += 1
                    continue
                if node_set_for[v] != -1:
                    assert node_set_for[v] == nsi
                    continue
                node_set_for[v] = nsi
                ns.append(int(v))
        node_sets.append(ns)
    return node_sets, node_set_for, mismatches


def periodic_condition(mesh, eps=1e-7, ignore_mismatch=False, ignore_dims=()):
    """PeriodicCondition (BoundaryConditions.hh:457-561).  Returns (dof_for_node, num_dofs, is_periodic_be)."""
    N = mesh.N
    bn = mesh.bdry_nodes
    P = mesh.nodes[bn]
    lo, hi = mesh.bbox_min, mesh.bbox_max
    on_min = np.abs(P - lo) <= eps              # FaceMembership (:44-50)
    on_max = np.abs(P - hi) <= eps
    if len(ignore_dims):
        # nodes on a periodic face lose the ignored faces; every other node loses all memberships (:472-500)
        periodic_dims = [d for d in range(N) if d not in ignore_dims]
        significant = (on_min[:, periodic_dims] | on_max[:, periodic_dims]).any(axis=1)
        for d in range(N):
            drop = ~significant if d not in ignore_dims else np.ones(bn.size, bool)
            on_min[drop, d] = False; on_max[drop, d] = False
    if ignore_mismatch:
        node_sets, node_set_for, _ = _match_periodic_permitting_mismatch(P, lo, hi, on_min, on_max, eps)
    else:
        node_sets, node_set_for = _match_periodic(P, lo, hi, on_min, on_max, eps)
    # boundary elements with all nodes on one common cell face (:126-146)
    memb = np.concatenate([on_min, on_max], axis=1)
    bnode = mesh.bdry_node_of_node[mesh.bdry_elem_nodes]
    common = memb[bnode].all(axis=1)
    if (common.sum(axis=1) > 1).any():
        raise RuntimeError("Boundary element on more than one cell face.")
    is_periodic_be = common.any(axis=1)
    # DoF ids in node order (BoundaryConditions.hh:533-554)
    dof = np.full(mesh.num_nodes, -1, dtype=np.int64)
    nd = 0
    for n in range(mesh.num_nodes):
        if dof[n] >= 0:
            continue
        b = mesh.bdry_node_of_node[n]
        if b >= 0:
            for j in node_sets[node_set_for[b]]:
                dof[bn[j]] = nd
        else:
            dof[n] = nd
        nd += 1
    return dof, nd, is_periodic_be


def periodic_condition_from_pairs(mesh, pairs):
    """PeriodicCondition(mesh, pcFile) (:563-610): DoFs = connected components of the identified-pair graph, numbered
    by their lowest node; no boundary element is marked periodic."""
    adj = [[] for _ in range(mesh.num_nodes)]
    for a, b in pairs:
        adj[a].append(b); adj[b].append(a)
    dof = np.full(mesh.num_nodes, -1, dtype=np.int64)
    nd = 0
    for n in range(mesh.num_nodes):
        if dof[n] >= 0:
            continue
        dof[n] = nd
        queue = [n]
        head = 0
        while head < len(queue):
            for v in adj[queue[head]]:
                if dof[v] < 0:
                    dof[v] = nd
                    queue.append(v)
            head += 1
        nd += 1
    return dof, nd, np.zeros(mesh.bdry_elem_nodes.shape[0], dtype=bool)


def canonical_basis(N, i):
    """SymmetricMatrix::CanonicalBasis (SymmetricMatrix.hh:407-413), flattened."""
    e = np.zeros(flat_len(N))
    e[i] = 1.0 if i < N else 0.5
    return e


def solve_cell_problems(sim):
    """PeriodicHomogenization::solveCellProblems (:34-54): mutates sim."""
    dof, nd, is_pbe = periodic_condition(sim.mesh)
    sim.set_periodic(dof, nd, is_pbe)
    sim.apply_no_rigid_motion_constraint()
    sim.set_use_pin_no_rigid_translation_constraint(True)
    w = []
    for i in range(flat_len(sim.N)):
        rhs = sim.constant_strain_load(-canonical_basis(sim.N, i))
        w.append(sim.solve(rhs))
    return w


def homogenized_tensor_displacement_form(sim, w_ij, base_cell_volume=0.0):
    """homogenizedElasticityTensorDisplacementForm (:146-186)."""
    m = sim.mesh; N = sim.N
    if base_cell_volume == 0.0:
        base_cell_volume = float(np.prod(m.bbox_max - m.bbox_min))
    F = flat_len(N)
    Eh = np.zeros((F, F))
    wts = integrated_phis(N - 1, m.deg)
    for i, w in enumerate(w_ij):
        w_int = np.einsum("n,b,bnc->bc", wts, m.bdry_vol, w[m.bdry_elem_nodes])      # int_be w dA
        nw = 0.5 * (w_int[:, :, None] * m.bdry_normal[:, None, :] + w_int[:, None, :] * m.bdry_normal[:, :, None])
        nw_flat = np.stack([nw[:, a, b] for a, b in (unflatten_index(N, k) for k in range(F))], axis=1)
        Eh[i] += double_contract(N, sim.D, nw_flat).sum(axis=0)
    Eh += sim.D * m.vol.sum()
    return Eh / base_cell_volume


def homogenized_tensor(sim, w_ij, base_cell_volume=0.0):
    """homogenizedElasticityTensor, stress-like volume form (:72-100)."""
    m = sim.mesh; N = sim.N
    if base_cell_volume == 0.0:
        base_cell_volume = float(np.prod(m.bbox_max - m.bbox_min))
    F = flat_len(N)
    Eh = np.zeros((F, F))
    for i, w in enumerate(w_ij):
        strain, stress = average_strain_stress(m, sim.D, w)
        Eh[i] += (stress * m.vol[:, None]).sum(axis=0)
    Eh += sim.D * m.vol.sum()
    return Eh / base_cell_volume


# ----------------------------------------------------------------------------
# Gmsh 2.2 reader (MeshIO.cc:625-760) -- fixtures only
# ----------------------------------------------------------------------------
def set_node_positions(mesh, vertices):
    """FEMMesh::setNodePositions (FEMMesh.hh:221-237): re-embed with new vertex positions, edge nodes at the
    edge midpoints; connectivity and numbering untouched."""
    N = mesh.N
    V = np.asarray(vertices, dtype=float)[:, :N].copy()
    nV = mesh.simplices.max() + 1 if mesh.deg == 1 else mesh.vertices.shape[0]
    assert V.shape[0] == mesh.vertices.shape[0]
    S = mesh.simplices
    if mesh.deg == 2:
        # edge node k sits between the two vertices of the first (element, local edge) that references it
        nedge = num_edges(N)
        mid = np.zeros((mesh.nodes.shape[0] - V.shape[0], N))
        for ei in range(nedge):
            ids = mesh.elem_nodes[:, N + 1 + ei] - V.shape[0]
            mid[ids] = 0.5 * (V[S[:, EDGE_START[ei]]] + V[S[:, EDGE_END[ei]]])
        mesh.nodes = np.vstack([V, mid])
    else:
        mesh.nodes = V.copy()
    mesh.vertices = V
    mesh.vol, mesh.G = embed_simplices(V[S])
    if mesh.bdry_elem_vertices.shape[0]:
        mesh.bdry_vol, mesh.bdry_normal = embed_boundary(V[mesh.bdry_elem_vertices])
    mesh.bbox_min = mesh.nodes.min(axis=0)
    mesh.bbox_max = mesh.nodes.max(axis=0)
    del nV
    return mesh


def transform_tensor(N, D, R):
    """ElasticityTensor::transform (ElasticityTensor.hh:515-541): E'_ijkl = E_pqrs R_ip R_jq R_kr R_ls."""
    C = tensor_C(N, D)
    R = np.asarray(R, dtype=float)
    Ct = np.einsum("pqrs,ip,jq,kr,ls->ijkl", C, R, R, R, R)
    F = flat_len(N)
    out = np.zeros((F, F))
    for a in range(F):
        i, j = unflatten_index(N, a)
        for b in range(F):
            k, l = unflatten_index(N, b)
            out[a, b] = Ct[i, j, k, l]
    return out


def frobenius_norm_sq(N, D):
    """ElasticityTensor::frobeniusNormSq = quadrupleContract(self) (:498-508): sum over all ijkl."""
    C = tensor_C(N, D)
    return float((C * C).sum())


def deformed_cell_homogenization(N, deg, vertices, simplices, D, jacobian, transform_version=False):
    """DeformedCells_cli --homogenize (src/bin/DeformedCells_cli.cc:222-262).  Direct version: periodic conditions
    matched on the undeformed cell, nodes moved by x -> J (x - centre), Eh over |bbox| det J.  Transform version:
    undeformed cell with the pulled-back material E.transform(J^-1), result pushed forward with J."""
    J = np.asarray(jacobian, dtype=float)
    sim = Simulator(N, deg, vertices, simplices)
    if transform_version:
        sim.set_material(transform_tensor(N, D, np.linalg.inv(J)))
        w = solve_cell_problems(sim)
        return transform_tensor(N, homogenized_tensor_displacement_form(sim, w), J), w, sim
    sim.set_material(D)
    m = sim.mesh
    bbox_volume = float(np.prod(m.bbox_max - m.bbox_min))
    center = 0.5 * (m.bbox_min + m.bbox_max)
    dof, nd, is_pbe = periodic_condition(m)
    sim.set_periodic(dof, nd, is_pbe)
    sim.apply_no_rigid_motion_constraint()
    sim.set_use_pin_no_rigid_translation_constraint(True)
    set_node_positions(m, (m.vertices - center) @ J.T)
    if hasattr(sim, "_K"):
        del sim._K
    w = []
    for i in range(flat_len(N)):
        w.append(sim.solve(sim.constant_strain_load(-canonical_basis(N, i))))
    return homogenized_tensor_displacement_form(sim, w, bbox_volume * float(np.linalg.det(J))), w, sim


def macro_to_micro_strain_tensors(sim, w_ij):
    """macroStrainToMicroStrainTensors (PeriodicHomogenization.hh:188-210): per element the flattened (F x F, no major
    symmetry) tensor whose column kl is avg strain(w_kl) + e_kl."""
    N = sim.N; F = flat_len(N)
    G = np.zeros((sim.mesh.num_elements, F, F))
    for kl, w in enumerate(w_ij):
        strain, _ = average_strain_stress(sim.mesh, sim.D, w)
        G[:, :, kl] = strain + canonical_basis(N, kl)[None, :]
    return G


def macro_to_micro_stress_tensors(sim, w_ij, Eh):
    """PeriodicHomogenization_cli --m2mstress (:173-186): E : G_e : Eh^-1 with F(A : B) = F(A) S F(B), S the shear
    doubler (ElasticityTensor.hh:483-495).  Returns the flattened (numElements, F, F) tensors."""
    N = sim.N
    dbl = np.ones(flat_len(N)); dbl[N:] = 2.0
    Sh = np.linalg.inv(Eh) / np.outer(dbl, dbl)           # ElasticityTensor::inverse (:315-323)
    G = macro_to_micro_strain_tensors(sim, w_ij)
    GS = np.einsum("eik,k,kj->eij", G, dbl, Sh)
    return np.einsum("ik,k,ekj->eij", sim.D, dbl, GS)


def unflatten_rank4(N, Mflat):
    """(i,j,k,l) array of a flattened rank-4 tensor with the minor symmetries."""
    out = np.zeros((N,) * 4)
    for i in range(N):
        for j in range(N):
            for k in range(N):
                for l in range(N):
                    out[i, j, k, l] = Mflat[flatten_indices(N, i, j), flatten_indices(N, k, l)]
    return out


# ----------------------------------------------------------------------------
# Discrete shape derivatives (LinearElasticity.hh:232-331, 1286-1373; EmbeddedElement.hh:269-372;
# PeriodicHomogenization.hh:383-563).  Vertex perturbations delta_p move the straight-sided elements; nodal values
# are transported (Lagrangian derivative).  With the piecewise-linear velocity field dp_h = sum_k delta_p_k lambda_k:
#   delta grad lambda_i = -(grad dp_h)^T grad lambda_i          (EmbeddedElement.hh:269-278)
#   delta vol / vol     = div dp_h                               (:365-372)
# and the same rule for every grad phi_i, whose barycentric coefficients do not move (:338-363).
# ----------------------------------------------------------------------------
def _shape_derivative_context(mesh, delta_p):
    N, deg = mesh.N, mesh.deg
    T = grad_phi_interpolant(N, deg)
    P, w = quadrature_points(N, 2 * (deg - 1))
    Tq = T[:, 0, :][None] if deg == 1 else np.einsum("qv,iva->qia", P, T)
    gp = np.einsum("qia,era->eqir", Tq, mesh.G)                       # grad phi_i at the quadrature points
    gavg = np.einsum("ia,era->eir", T.mean(axis=1), mesh.G)           # int grad phi_i / vol
    grad_dp = None
    if delta_p is not None:
        dpe = np.asarray(delta_p, dtype=float)[mesh.simplices]        # (ne, N+1, N)
        grad_dp = np.einsum("eka,erk->ear", dpe, mesh.G)              # d(dp_a)/dx_r, constant per element
    return gp, gavg, w, grad_dp


def _element_C(N, D, ne):
    D = np.asarray(D)
    if D.ndim == 2:
        return np.broadcast_to(tensor_C(N, D), (ne,) + (N,) * 4)
    return np.stack([tensor_C(N, D[e]) for e in range(ne)])


def apply_delta_stiffness_matrix(mesh, D, u_nodes, delta_p, dof_for_node=None, num_dofs=None):
    """Simulator::applyDeltaStiffnessMatrix (:1301-1330): (delta K) u for the fixed per-NODE field u under the vertex
    perturbation delta_p, as a per-DoF load."""
    N = mesh.N
    gp, _, w, gdp = _shape_derivative_context(mesh, delta_p)
    C = _element_C(N, D, mesh.num_elements)
    div = np.einsum("eaa->e", gdp)
    ue = np.asarray(u_nodes)[mesh.elem_nodes]
    gu = np.einsum("eic,eqir->eqcr", ue, gp)
    dgu = -np.einsum("eqcm,emr->eqcr", gu, gdp)
    sym = lambda g: 0.5 * (g + np.swapaxes(g, -1, -2))
    sig = np.einsum("eabcd,eqcd->eqab", C, sym(gu))
    dsig = np.einsum("eabcd,eqcd->eqab", C, sym(dgu))
    dgp = -np.einsum("emr,eqim->eqir", gdp, gp)
    f = (np.einsum("e,eqcr,eqir->eqic", div, sig, gp) + np.einsum("eqcr,eqir->eqic", dsig, gp)
         + np.einsum("eqcr,eqir->eqic", sig, dgp))
    f = np.einsum("q,e,eqic->eic", w, mesh.vol, f)
    nd = mesh.num_nodes if dof_for_node is None else num_dofs
    dof = mesh.elem_nodes if dof_for_node is None else np.asarray(dof_for_node)[mesh.elem_nodes]
    load = np.zeros((nd, N))
    np.add.at(load, dof.reshape(-1), f.reshape(-1, N))
    return load


def delta_constant_strain_load(mesh, D, strain_flat, delta_p, dof_for_node=None, num_dofs=None):
    """Simulator::deltaConstantStrainLoad (:1333-1348) with deltaPerElementConstantStrainLoad (:286-302)."""
    N = mesh.N
    _, gavg, _, gdp = _shape_derivative_context(mesh, delta_p)
    C = _element_C(N, D, mesh.num_elements)
    s = np.einsum("eabcd,cd->eab", C, flat_to_sym(N, np.asarray(strain_flat, dtype=float)))
    div = np.einsum("eaa->e", gdp)
    dgavg = -np.einsum("emr,eim->eir", gdp, gavg)
    l = np.einsum("e,ecr,eir->eic", mesh.vol * div, s, gavg) + np.einsum("e,ecr,eir->eic", mesh.vol, s, dgavg)
    nd = mesh.num_nodes if dof_for_node is None else num_dofs
    dof = mesh.elem_nodes if dof_for_node is None else np.asarray(dof_for_node)[mesh.elem_nodes]
    load = np.zeros((nd, N))
    np.add.at(load, dof.reshape(-1), l.reshape(-1, N))
    return load


def delta_average_strain_field(mesh, u_nodes, delta_u_nodes, delta_p):
    """Simulator::deltaAverageStrainField (:1365-1375): avg (delta strain)(u) + avg strain(delta u), flattened."""
    N = mesh.N
    _, gavg, _, gdp = _shape_derivative_context(mesh, delta_p)
    gu = np.einsum("eic,eir->ecr", np.asarray(u_nodes)[mesh.elem_nodes], gavg)
    gdu = np.einsum("eic,eir->ecr", np.asarray(delta_u_nodes)[mesh.elem_nodes], gavg)
    g = -np.einsum("ecm,emr->ecr", gu, gdp) + gdu
    eps = 0.5 * (g + np.swapaxes(g, 1, 2))
    return np.stack([eps[:, a, b] for a, b in (unflatten_index(N, i) for i in range(flat_len(N)))], axis=1)


def homogenized_tensor_discrete_differential(sim, w_ij):
    """homogenizedElasticityTensorDiscreteDifferential (PeriodicHomogenization.hh:383-478): dCh[v, c] (flattened
    F x F) such that delta Ch = sum_v sum_c dCh[v, c] delta_p[v, c]; fluctuations held fixed (they are stationary
    points of the cell energy, so this is the full derivative)."""
    m = sim.mesh; N = sim.N; F = flat_len(N)
    gp, _, wq, _ = _shape_derivative_context(m, None)
    C = _element_C(N, sim.D, m.num_elements)
    gw = np.stack([np.einsum("eic,eqir->eqcr", np.asarray(w)[m.elem_nodes], gp) for w in w_ij])      # (F, ne, q, N, N)
    eps = 0.5 * (gw + np.swapaxes(gw, -1, -2))
    for ij in range(F):
        eps[ij] += flat_to_sym(N, canonical_basis(N, ij))
    sig = np.einsum("eabcd,feqcd->feqab", C, eps)
    energy = np.einsum("feqab,geqab->eqfg", eps, sig)
    wv = np.einsum("q,e->eq", wq, m.vol)
    t1 = np.einsum("eq,eqfg,ecv->evcfg", wv, energy, m.G)
    A = np.einsum("feqac,geqab->eqfgcb", gw, sig)
    A = A + np.swapaxes(A, 2, 3)                                       # grad w_ij^T sig_kl + grad w_kl^T sig_ij
    t2 = np.einsum("eq,eqfgcb,ebv->evcfg", wv, A, m.G)
    out = np.zeros((m.vertices.shape[0], N, F, F))
    np.add.at(out, m.simplices.reshape(-1), (t1 - t2).reshape(-1, N, F, F))
    return out / float(np.prod(m.bbox_max - m.bbox_min))


def delta_fluctuation_displacements(sim, w_ij, delta_p):
    """deltaFluctuationDisplacements (:520-540): K dw_ij = delta load(-e_ij) - (delta K) w_ij with the cell problem's
    constraints (periodic DoFs, pinned node at value 0)."""
    out = []
    N = sim.N
    K = sim.stiffness()
    fixed, vals = sim.fixed_vars_and_values()
    for ij, w in enumerate(w_ij):
        rhs = delta_constant_strain_load(sim.mesh, sim.D, -canonical_basis(N, ij), delta_p, sim.dof_for_node, sim.num_dofs())
        rhs = rhs - apply_delta_stiffness_matrix(sim.mesh, sim.D, w, delta_p, sim.dof_for_node, sim.num_dofs())
        x = solve_fixed(K, rhs.reshape(-1), fixed, np.zeros_like(vals))
        out.append(sim.dof_to_node_field(x))
    return out


def read_msh(path):
    """Returns (vertices (n,3), elements (m,k), gmsh element type).  ASCII or binary, 8-byte reals."""
    npe = {2: 3, 4: 4, 3: 4, 5: 8, 9: 6, 11: 10, 1: 2, 8: 3}
    with open(path, "rb") as f:
        data = f.read()
    pos = 0

    def line():
        nonlocal pos
        while True:
            e = data.index(b"\n", pos)
            s = data[pos:e].decode("latin1").strip()
            pos = e + 1
            if s and not s.startswith("#"):
                return s
    assert line() == "$MeshFormat"
    ver, ftype, dsize = line().split()
    binary = int(ftype) == 1
    assert int(dsize) == 8
    if binary:
        one = struct.unpack_from("<i", data, pos)[0]; pos += 4
        assert one == 1
    assert line() == "$EndMeshFormat"
    assert line() == "$Nodes"
    nn = int(line())
    V = np.zeros((nn, 3))
    if binary:
        rec = np.frombuffer(data, dtype=np.dtype([("id", "<i4"), ("p", "<f8", 3)]), count=nn, offset=pos)
        assert np.array_equal(rec["id"], np.arange(1, nn + 1))
        V[:] = rec["p"]; pos += nn * 28
    else:
        for i in range(nn):
            t = line().split(); assert int(t[0]) == i + 1
            V[i] = [float(x) for x in t[1:4]]
    assert line() == "$EndNodes"
    assert line() == "$Elements"
    ne = int(line())
    E = None; etype = None
    if binary:
        read = 0
        rows = []
        while read < ne:
            et, cnt, ntags = struct.unpack_from("<3i", data, pos); pos += 12
            k = npe[et]; etype = et
            w = 1 + ntags + k
            blk = np.frombuffer(data, dtype="<i4", count=cnt * w, offset=pos).reshape(cnt, w); pos += cnt * w * 4
            rows.append(blk[:, 1 + ntags:] - 1); read += cnt
        E = np.vstack(rows)
    else:
        rows = []
        for i in range(ne):
            t = [int(x) for x in line().split()]
            etype = t[1]; ntags = t[2]
            rows.append([x - 1 for x in t[3 + ntags:3 + ntags + npe[etype]]])
        E = np.array(rows)
    return V, E.astype(np.int64), etype


# ---------------------------------------------------------------------------------------------
# Orthotropic-cell homogenization (OrthotropicHomogenization.hh:42-240): the positive octant
# (quadrant) of a base cell with reflective symmetries; no periodic DoFs, one K, 1 + (flat - N)
# different sets of fixed variables.
def orthotropic_fixed_variable_sets(mesh, eps=1e-7):
    """[stretch set, shear set 0, ...] of scalar variable indices (:77-119)."""
    N = mesh.N
    bn = mesh.bdry_nodes
    P = mesh.nodes[bn]
    on = (np.abs(P - mesh.bbox_min[None, :]) <= eps) | (np.abs(P - mesh.bbox_max[None, :]) <= eps)   # (nbn, N): onMinOrMaxFace(c)
    stretch = sorted(int(N * n + c) for k, n in enumerate(bn) for c in range(N) if on[k, c])
    sets = [stretch]
    for s in range(flat_len(N) - N):
        fix = set()
        for k, n in enumerate(bn):
            for c in range(N):
                if on[k, c]:
                    if N == 3:
                        fix.add(int(N * n + s))
                        if c != s:
                            fix.add(int(N * n + (N - (c + s))))
                    else:
                        fix.add(int(N * n + (1 if c == 0 else 0)))
        sets.append(sorted(fix))
    return sets


def solve_orthotropic_cell_problems(sim, eps=1e-7):
    """Orthotropic::solveCellProblems (:42-145): w_ij on the orthotropic base cell."""
    N = sim.N
    F = flat_len(N)
    K = stiffness_matrix(sim.mesh, sim.D)
    sets = orthotropic_fixed_variable_sets(sim.mesh, eps)
    w = []
    for ij in range(F):
        rhs = constant_strain_load(sim.mesh, sim.D, -canonical_basis(N, ij)).reshape(-1)
        fixed = sets[0] if ij < N else sets[ij - N + 1]
        w.append(solve_fixed(K, rhs, np.array(fixed, dtype=np.int64), np.zeros(len(fixed))).reshape(-1, N))
    return w


def fluctuation_displacement_sign(N, ij, r):
    """:149-163"""
    if ij < N:
        return 1.0
    bits = [(r >> b) & 1 for b in range(N)]
    if N == 3:
        bits[ij - N] = 0
    return -1.0 if sum(bits) == 1 else 1.0


def homogenized_tensor_from_ortho_cell_quantity(N, EhO):
    """:165-185 -- only the upper triangle of EhO is read; the result is symmetric."""
    F = flat_len(N)
    Eh = np.zeros((F, F))
    for r in range(1 << N):
        for kl in range(F):
            for ij in range(kl + 1):
                Eh[ij, kl] += fluctuation_displacement_sign(N, ij, r) * fluctuation_displacement_sign(N, kl, r) * EhO[ij, kl]
    Eh /= (1 << N)
    return np.triu(Eh) + np.triu(Eh, 1).T


def orthotropic_homogenized_tensor_displacement_form(sim, w_ij):
    return homogenized_tensor_from_ortho_cell_quantity(sim.N, homogenized_tensor_displacement_form(sim, w_ij))
