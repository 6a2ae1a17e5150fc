"""ctypes access to libmeshfem_host.so: the host C++ mirror of the reference's mesh layer
(FEMMesh numbering, `grid -t` generator, MeshIO).  Used by tests and bench.py to build inputs;
none of this touches the GPU."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p
from types import SimpleNamespace

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmeshfem_host.so")
_lib = None


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: run `python -m meshfem_b200.build`")
        lib = ctypes.CDLL(LIB_PATH)
        lib.mfemhost_last_error.restype = c_char_p
        lib.mfemhost_grid.restype = c_void_p
        lib.mfemhost_grid.argtypes = [c_int, POINTER(c_int64), POINTER(c_double), POINTER(c_double)]
        lib.mfemhost_load_mesh.restype = c_void_p
        lib.mfemhost_load_mesh.argtypes = [c_char_p, POINTER(c_int)]
        lib.mfemhost_from_arrays.restype = c_void_p
        lib.mfemhost_from_arrays.argtypes = [c_int, c_int64, POINTER(c_double), c_int64, POINTER(c_int64)]
        lib.mfemhost_free.argtypes = [c_void_p]
        lib.mfemhost_raw_sizes.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64)]
        lib.mfemhost_raw_copy.argtypes = [c_void_p, POINTER(c_double), POINTER(c_int64)]
        lib.mfemhost_build_femmesh.argtypes = [c_void_p, c_int, POINTER(c_int64)]
        lib.mfemhost_femmesh_copy.argtypes = [c_void_p, POINTER(c_double), POINTER(c_int32), POINTER(c_int32),
                                              POINTER(c_int32), POINTER(c_int32), POINTER(c_double),
                                              POINTER(c_double), POINTER(c_double)]
        lib.mfemhost_apply_bc.argtypes = [c_void_p, c_int, c_char_p, c_int, POINTER(c_int64)]
        lib.mfemhost_bc_rows.argtypes = [POINTER(c_int64), POINTER(c_double), POINTER(c_double), POINTER(c_double)]
        lib.mfemhost_bc_copy.argtypes = [POINTER(c_int64), POINTER(c_double), POINTER(c_double), POINTER(c_int64),
                                         ctypes.POINTER(ctypes.c_uint8)]
        lib.mfemhost_material.argtypes = [c_int, c_char_p, POINTER(c_double), ctypes.c_char_p, c_int]
        lib.mfemhost_eval_expr.argtypes = [c_char_p, c_double, c_double, c_double, POINTER(c_double)]
        lib.mfemhost_save_mesh.argtypes = [c_void_p, c_char_p]
        lib.mfemhost_perforated_cell.argtypes = [c_int, c_int64, c_int64]
        lib.mfemhost_strain_field.argtypes = [c_void_p, c_int, POINTER(c_double), c_int, POINTER(c_double), POINTER(c_double),
                                              c_char_p, c_int]
        lib.mfemhost_perforated_cell.restype = c_void_p
        lib.mfemhost_msh_field.argtypes = [c_int, c_char_p, c_char_p, c_int, c_int, POINTER(c_double), c_int64,
                                           POINTER(c_int64), POINTER(c_int)]
        lib.mfemhost_tensor_analysis.argtypes = [c_int, POINTER(c_double)] + [POINTER(c_double)] * 5
        lib.mfemhost_closest_isotropic.argtypes = [c_int, POINTER(c_double), POINTER(c_double)]
        lib.mfemhost_partition.argtypes = [c_int, c_int64, POINTER(c_double), c_int64, c_int, POINTER(c_int32), c_int, c_int,
                                           POINTER(c_int64), c_int64, POINTER(c_int64)]
        lib.mfemhost_set_partitioner.argtypes = [c_int]
        lib.mfemhost_partition_copy.argtypes = [POINTER(c_int64), POINTER(c_int64), POINTER(c_int32),
                                                ctypes.POINTER(ctypes.c_uint8), POINTER(c_int32), POINTER(c_int64),
                                                POINTER(c_int32), POINTER(c_int64), POINTER(c_int64)]
        _lib = lib
    return _lib


def _err(lib):
    return RuntimeError(lib.mfemhost_last_error().decode())


class RawMesh:
    """Vertices + simplices as read / generated (MeshIO::IOVertex / IOElement)."""

    def __init__(self, ptr, dim):
        self._p, self.dim = ptr, dim
        self.lib = load_library()

    def __del__(self):
        if getattr(self, "_p", None):
            self.lib.mfemhost_free(self._p)
            self._p = None

    def arrays(self):
        nV, nE = c_int64(), c_int64()
        self.lib.mfemhost_raw_sizes(self._p, ctypes.byref(nV), ctypes.byref(nE))
        V = np.zeros((nV.value, 3)); E = np.zeros((nE.value, self.dim + 1), dtype=np.int64)
        self.lib.mfemhost_raw_copy(self._p, V.ctypes.data_as(POINTER(c_double)), E.ctypes.data_as(POINTER(c_int64)))
        return V, E

    def save(self, path):
        if self.lib.mfemhost_save_mesh(self._p, os.fsencode(path)) != 0:
            raise _err(self.lib)

    def apply_bc(self, deg, bc_json_text, periodic=False, pin=None):
        """Host-side Simulator bookkeeping (host-only mode): returns dict(fixed_vars, fixed_vals,
        load[numDoFs, dim], dof_for_node, num_dofs, internal_be) and, for configurations with Lagrange-multiplier
        rows (no_rigid_motion, unconstrained translations; periodic with pin=False; pin = the
        setUsePinNoRigidTranslationConstraint option, default on under periodicity and off otherwise), constraint_rows [m, N*numDoFs],
        constraint_rhs [m] and the candidate rigid_modes of the null space."""
        sz = (c_int64 * 3)()
        if self.lib.mfemhost_apply_bc(self._p, deg, (bc_json_text or "").encode(), ((2 if pin is False else 1) if periodic else (4 if pin else 0)), sz) != 0:
            raise _err(self.lib)
        nfix, ndof, nbe = (int(x) for x in sz)
        V, E = self.arrays()
        fm = self.femmesh(deg)
        fixed = np.zeros(nfix, dtype=np.int64); vals = np.zeros(nfix); load = np.zeros((ndof, self.dim))
        dof = np.zeros(fm.num_nodes, dtype=np.int64); ibe = np.zeros(nbe, dtype=np.uint8)
        self.lib.mfemhost_bc_copy(fixed.ctypes.data_as(POINTER(c_int64)), vals.ctypes.data_as(POINTER(c_double)),
                                  load.ctypes.data_as(POINTER(c_double)), dof.ctypes.data_as(POINTER(c_int64)),
                                  ibe.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        cnt = (c_int64 * 3)()
        self.lib.mfemhost_bc_rows(cnt, None, None, None)
        nrows, nmodes, nvar = (int(x) for x in cnt)
        rows = np.zeros((nrows, nvar)); rhs = np.zeros(nrows); modes = np.zeros((nmodes, nvar))
        if nrows:
            dp = POINTER(c_double)
            self.lib.mfemhost_bc_rows(cnt, rows.ctypes.data_as(dp), rhs.ctypes.data_as(dp), modes.ctypes.data_as(dp))
        return dict(fixed_vars=fixed, fixed_vals=vals, load=load, dof_for_node=dof, num_dofs=ndof,
                    internal_be=ibe.astype(bool), mesh=fm, constraint_rows=rows, constraint_rhs=rhs, rigid_modes=modes)

    def strain_field(self, deg, u_nodes, D=None, stress=False, path=None, binary=True):
        """Simulator::strainField / stressField (host): (numElements, nodesPerElem, flat) nodal values of the
        upsampled interpolant; path: also write u + the field as a full-degree .msh ($ElementNodeData)."""
        K = self.dim
        F = K * (K + 1) // 2
        npe = K + 1 if deg == 1 else (6 if K == 2 else 10)
        _, E = self.arrays()
        u = np.ascontiguousarray(u_nodes, dtype=np.float64)
        Dm = np.ascontiguousarray(np.eye(F) if D is None else D, dtype=np.float64)
        out = np.zeros((E.shape[0], npe, F))
        dp = POINTER(c_double)
        if self.lib.mfemhost_strain_field(self._p, deg, u.ctypes.data_as(dp), 1 if stress else 0, Dm.ctypes.data_as(dp),
                                          out.ctypes.data_as(dp), None if path is None else os.fsencode(path), 1 if binary else 0) != 0:
            raise _err(self.lib)
        return out

    def deformed_displacement_form(self, deg, D, J, w_ij):
        """Host half of DeformedCells_cli --homogenize: (Eh, deformed node positions) for fluctuation displacements
        w_ij[flat, numNodes, dim] on the cell deformed by x -> J (x - centre); J None = undeformed."""
        K = self.dim
        F = K * (K + 1) // 2
        dp = POINTER(c_double)
        fm = self.femmesh(deg)
        w = np.ascontiguousarray(w_ij, dtype=np.float64)
        assert w.shape == (F, fm.num_nodes, K)
        Dm = np.ascontiguousarray(D, dtype=np.float64)
        Jm = None if J is None else np.ascontiguousarray(J, dtype=np.float64)
        Eh = np.zeros((F, F)); nodes = np.zeros((fm.num_nodes, K))
        self.lib.mfemhost_deformed_displacement_form.argtypes = [c_void_p, c_int, dp, dp, dp, dp, dp]
        if self.lib.mfemhost_deformed_displacement_form(self._p, deg, Dm.ctypes.data_as(dp), None if Jm is None else Jm.ctypes.data_as(dp),
                                                        w.ctypes.data_as(dp), Eh.ctypes.data_as(dp), nodes.ctypes.data_as(dp)) != 0:
            raise _err(self.lib)
        return Eh, nodes

    def periodic_condition(self, deg, eps=1e-7, ignore_mismatch=False, ignore_dims=(), pairs_file=None):
        """PeriodicCondition of the reference: (dof_for_node, num_dofs, is_periodic_be)."""
        fm = self.femmesh(deg)
        dof = np.zeros(fm.num_nodes, dtype=np.int64)
        ibe = np.zeros(fm.bdry_elem_nodes.shape[0], dtype=np.uint8)
        nd = c_int64()
        ig = np.ascontiguousarray(ignore_dims, dtype=np.int64)
        self.lib.mfemhost_periodic_condition.argtypes = [c_void_p, c_int, c_double, c_int, c_int, POINTER(c_int64), c_char_p,
                                                         POINTER(c_int64), POINTER(ctypes.c_uint8), POINTER(c_int64)]
        if self.lib.mfemhost_periodic_condition(self._p, deg, eps, 1 if ignore_mismatch else 0, ig.size, ig.ctypes.data_as(POINTER(c_int64)),
                                                None if pairs_file is None else os.fsencode(pairs_file),
                                                dof.ctypes.data_as(POINTER(c_int64)), ibe.ctypes.data_as(POINTER(ctypes.c_uint8)),
                                                ctypes.byref(nd)) != 0:
            raise _err(self.lib)
        return dof, int(nd.value), ibe.astype(bool)

    def shape_derivatives(self, deg, D, u, du, strain, delta_p, w_ij=None, periodic=False, num_dofs=None):
        """Discrete shape derivatives of the host Simulator for the per-vertex perturbation delta_p: dict with
        dKu (applyDeltaStiffnessMatrix), dload (deltaConstantStrainLoad), dstrain (deltaAverageStrainField) and, when
        w_ij is given, dCh[numVertices, dim, F, F] (homogenizedElasticityTensorDiscreteDifferential)."""
        K = self.dim
        F = K * (K + 1) // 2
        dp_ = POINTER(c_double)
        fm = self.femmesh(deg)
        nd = fm.num_nodes if num_dofs is None else num_dofs
        arr = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        D, u, du, strain, delta_p = arr(D), arr(u), arr(du), arr(strain), arr(delta_p)
        assert u.shape == (fm.num_nodes, K) and du.shape == u.shape and delta_p.shape == (fm.num_vertices, K)
        dKu = np.zeros((nd, K)); dload = np.zeros((nd, K)); dstrain = np.zeros((fm.num_elements, F))
        w = None if w_ij is None else arr(w_ij)
        dCh = None if w is None else np.zeros((fm.num_vertices, K, F, F))
        self.lib.mfemhost_shape_derivatives.argtypes = [c_void_p, c_int, c_int] + [dp_] * 10
        if self.lib.mfemhost_shape_derivatives(self._p, deg, 1 if periodic else 0, D.ctypes.data_as(dp_), u.ctypes.data_as(dp_),
                                               du.ctypes.data_as(dp_), strain.ctypes.data_as(dp_), delta_p.ctypes.data_as(dp_),
                                               None if w is None else w.ctypes.data_as(dp_), dKu.ctypes.data_as(dp_),
                                               dload.ctypes.data_as(dp_), dstrain.ctypes.data_as(dp_),
                                               None if dCh is None else dCh.ctypes.data_as(dp_)) != 0:
            raise _err(self.lib)
        return dict(dKu=dKu, dload=dload, dstrain=dstrain, dCh=dCh)

    def femmesh(self, deg):
        """FEMMesh<dim,deg> flat data with the reference's numbering."""
        sz = (c_int64 * 5)()
        if self.lib.mfemhost_build_femmesh(self._p, deg, sz) != 0:
            raise _err(self.lib)
        nn, ne, nbe, nbn, nv = (int(x) for x in sz)
        K = self.dim
        npe = K + 1 if deg == 1 else (6 if K == 2 else 10)
        npbe = K if deg == 1 else (3 if K == 2 else 6)
        m = SimpleNamespace(N=K, deg=deg, num_nodes=nn, num_elements=ne, num_vertices=nv)
        m.nodes = np.zeros((nn, K)); m.elem_nodes = np.zeros((ne, npe), dtype=np.int32)
        m.bdry_elem_nodes = np.zeros((nbe, npbe), dtype=np.int32); m.bdry_elem_vertices = np.zeros((nbe, K), dtype=np.int32)
        m.bdry_nodes = np.zeros(nbn, dtype=np.int32); m.bdry_vol = np.zeros(nbe); m.bdry_normal = np.zeros((nbe, K))
        bbox = np.zeros(6)
        dp, ip = POINTER(c_double), POINTER(c_int32)
        self.lib.mfemhost_femmesh_copy(self._p, m.nodes.ctypes.data_as(dp), m.elem_nodes.ctypes.data_as(ip),
                                       m.bdry_elem_nodes.ctypes.data_as(ip), m.bdry_elem_vertices.ctypes.data_as(ip),
                                       m.bdry_nodes.ctypes.data_as(ip), m.bdry_vol.ctypes.data_as(dp),
                                       m.bdry_normal.ctypes.data_as(dp), bbox.ctypes.data_as(dp))
        m.bbox_min, m.bbox_max = bbox[:K].copy(), bbox[3:3 + K].copy()
        return m


def material_tensor(dim, json_text):
    """Materials::Constant<dim>::setFromJson -> flattened tensor, plus its anisotropic round-trip JSON."""
    lib = load_library()
    F = dim * (dim + 1) // 2
    D = np.zeros((F, F))
    buf = ctypes.create_string_buffer(8192)
    if lib.mfemhost_material(dim, json_text.encode(), D.ctypes.data_as(POINTER(c_double)), buf, 8192) != 0:
        raise _err(lib)
    return D, buf.value.decode()


def msh_field(dim, path, name, kind="scalar", domain="any"):
    """MSHFieldParser<dim>(path).{scalar,vector,symmetricMatrix}Field(name, domain) -> (values, actual domain)."""
    lib = load_library()
    k = {"scalar": 0, "vector": 1, "matrix": 2}[kind]
    d = {"element": 0, "node": 1, "any": 2}[domain]
    n, got = c_int64(), c_int()
    if lib.mfemhost_msh_field(dim, path.encode(), name.encode(), k, d, None, 0, ctypes.byref(n), ctypes.byref(got)) != 0:
        raise _err(lib)
    width = {0: 1, 1: dim, 2: dim * (dim + 1) // 2}[k]
    out = np.zeros((n.value, width))
    if lib.mfemhost_msh_field(dim, path.encode(), name.encode(), k, d, out.ctypes.data_as(POINTER(c_double)), out.size,
                              ctypes.byref(n), ctypes.byref(got)) != 0:
        raise _err(lib)
    return (out[:, 0] if k == 0 else out), ("element" if got.value == 0 else "node")


def tensor_analysis(D):
    """Eigenstrains, compliance, orthotropic parameters and anisotropy of a flattened elasticity tensor
    (ElasticityTensor::computeEigenstrains / inverse / getOrthotropic* / anisotropy)."""
    lib = load_library()
    D = np.ascontiguousarray(D, dtype=np.float64)
    F = D.shape[0]
    dim = 3 if F == 6 else 2
    lam, strains, S = np.zeros(F), np.zeros((F, F)), np.zeros((F, F))
    ortho, aniso = np.zeros(9 if dim == 3 else 4), c_double()
    dp = lambda a: a.ctypes.data_as(POINTER(c_double))
    if lib.mfemhost_tensor_analysis(dim, dp(D), dp(lam), dp(strains), dp(S), dp(ortho), ctypes.byref(aniso)) != 0:
        raise _err(lib)
    return SimpleNamespace(lambdas=lam, strains=strains, compliance=S, orthotropic=ortho, anisotropy=aniso.value)


def closest_isotropic_tensor(D):
    """closestIsotropicTensor (TensorProjection.hh): Frobenius projection onto the isotropic tensors."""
    lib = load_library()
    D = np.ascontiguousarray(D, dtype=np.float64)
    out = np.zeros_like(D)
    if lib.mfemhost_closest_isotropic(3 if D.shape[0] == 6 else 2, D.ctypes.data_as(POINTER(c_double)),
                                      out.ctypes.data_as(POINTER(c_double))) != 0:
        raise _err(lib)
    return out


def eval_expression(expr, x=0.0, y=0.0, z=0.0):
    lib = load_library()
    out = c_double()
    if lib.mfemhost_eval_expr(expr.encode(), x, y, z, ctypes.byref(out)) != 0:
        raise _err(lib)
    return out.value


def grid(sizes, min_corner=None, max_corner=None) -> RawMesh:
    """`grid AxB[xC] -t` of the reference (src/bin/tools/grid.cc)."""
    lib = load_library()
    sz = (c_int64 * len(sizes))(*sizes)
    mn = mx = None
    if min_corner is not None:
        mn = (c_double * 3)(*(list(min_corner) + [0.0] * (3 - len(min_corner))))
        mx = (c_double * 3)(*(list(max_corner) + [0.0] * (3 - len(max_corner))))
    p = lib.mfemhost_grid(len(sizes), sz, mn, mx)
    if not p:
        raise _err(lib)
    return RawMesh(p, len(sizes))


def perforated_cell(ndim, n, hole) -> RawMesh:
    """BASELINE config-4 family: n^ndim voxels on the unit cell minus the centred hole^ndim block, tesselated."""
    lib = load_library()
    p = lib.mfemhost_perforated_cell(ndim, n, hole)
    if not p:
        raise _err(lib)
    return RawMesh(p, ndim)


def load_mesh(path) -> RawMesh:
    lib = load_library()
    dim = c_int()
    p = lib.mfemhost_load_mesh(os.fsencode(path), ctypes.byref(dim))
    if not p:
        raise _err(lib)
    return RawMesh(p, dim.value)


def from_arrays(dim, V, E) -> RawMesh:
    lib = load_library()
    V3 = np.zeros((len(V), 3)); V3[:, :np.asarray(V).shape[1]] = V
    E = np.ascontiguousarray(E, dtype=np.int64)
    p = lib.mfemhost_from_arrays(dim, V3.shape[0], V3.ctypes.data_as(POINTER(c_double)), E.shape[0],
                                 E.ctypes.data_as(POINTER(c_int64)))
    return RawMesh(p, dim)


def partition(m, n_parts, rank, dof_for_node=None, method=None):
    """Element partition of FEMMesh data `m` (from RawMesh.femmesh) -- method "slab" (default) or "rcb" (recursive
    coordinate bisection; None = slabs unless MESHFEM_PARTITIONER=rcb) --: this rank's local sub-mesh
    and interface description (include/MeshFEM/Partition.hh).  Returns a SimpleNamespace with
    elems, nodes_global, elem_nodes (local node ids), local node coordinates, and -- in terms of
    DoFs, which are the nodes unless `dof_for_node` (periodic identification) is given --
    dofs_global, dof_for_node (local, or None), owned mask, neighbor_ranks and shared[q] = local DoF
    ids shared with rank q (ascending global id)."""
    lib = load_library()
    assert method in (None, "slab", "rcb")
    lib.mfemhost_set_partitioner(-1 if method is None else (1 if method == "rcb" else 0))
    nodes = np.ascontiguousarray(m.nodes, dtype=np.float64)
    en = np.ascontiguousarray(m.elem_nodes, dtype=np.int32)
    sz = (c_int64 * 6)()
    dfn = None if dof_for_node is None else np.ascontiguousarray(dof_for_node, dtype=np.int64)
    n_dofs = 0 if dfn is None else int(dfn.max()) + 1
    if lib.mfemhost_partition(m.N, nodes.shape[0], nodes.ctypes.data_as(POINTER(c_double)), en.shape[0], en.shape[1],
                              en.ctypes.data_as(POINTER(c_int32)), n_parts, rank,
                              None if dfn is None else dfn.ctypes.data_as(POINTER(c_int64)), n_dofs, sz) != 0:
        raise _err(lib)
    ne, nn, nnb, nsh, nowned, nd = (int(x) for x in sz)
    p = SimpleNamespace(N=m.N, deg=m.deg, rank=rank, n_parts=n_parts, num_owned=nowned)
    p.elems = np.zeros(ne, dtype=np.int64); p.nodes_global = np.zeros(nn, dtype=np.int64)
    p.elem_nodes = np.zeros((ne, en.shape[1]), dtype=np.int32); p.owned = np.zeros(nd, dtype=np.uint8)
    p.neighbor_ranks = np.zeros(nnb, dtype=np.int32); p.neighbor_offsets = np.zeros(nnb + 1, dtype=np.int64)
    p.shared_local = np.zeros(nsh, dtype=np.int32)
    p.dofs_global = np.zeros(nd, dtype=np.int64)
    p.dof_for_node = None if dfn is None else np.zeros(nn, dtype=np.int64)
    ip, lp = POINTER(c_int32), POINTER(c_int64)
    lib.mfemhost_partition_copy(p.elems.ctypes.data_as(lp), p.nodes_global.ctypes.data_as(lp), p.elem_nodes.ctypes.data_as(ip),
                                p.owned.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), p.neighbor_ranks.ctypes.data_as(ip),
                                p.neighbor_offsets.ctypes.data_as(lp), p.shared_local.ctypes.data_as(ip),
                                p.dofs_global.ctypes.data_as(lp),
                                None if dfn is None else p.dof_for_node.ctypes.data_as(lp))
    p.nodes = nodes[p.nodes_global]
    p.num_nodes, p.num_elements, p.num_dofs = nn, ne, nd
    p.shared = {int(q): p.shared_local[p.neighbor_offsets[i]:p.neighbor_offsets[i + 1]] for i, q in enumerate(p.neighbor_ranks)}
    return p


def constrained_solve(rows, rows_rhs, fixed_vars, rigid_modes, fs, spsd_solve):
    """RigidMotionConstraints::solve (include/MeshFEM/RigidMotionConstraints.hh) -- the host algebra that resolves
    Lagrange-multiplier rows around an SPSD solver.  spsd_solve(B[nrhs, n]) -> U[nrhs, n] must solve
    K_ff u_f = b_f - K_fc u_c with the fixed values in place (the Simulator passes the device PCG; this entry
    exists so that the algebra can be exercised with any solver).  Returns (U[nrhs, n], multipliers[nrhs, m])."""
    lib = load_library()
    rows = np.ascontiguousarray(np.atleast_2d(rows), dtype=np.float64)
    m, n = rows.shape
    rows_rhs = np.ascontiguousarray(rows_rhs, dtype=np.float64)
    fixed_vars = np.ascontiguousarray(fixed_vars, dtype=np.int64)
    rigid_modes = np.ascontiguousarray(np.atleast_2d(rigid_modes), dtype=np.float64)
    fs = np.ascontiguousarray(np.atleast_2d(fs), dtype=np.float64)
    nrhs = fs.shape[0]
    us = np.zeros((nrhs, n)); lam = np.zeros((nrhs, m))
    CB = ctypes.CFUNCTYPE(c_int, c_int, POINTER(c_double), POINTER(c_double))
    failure = []

    def _cb(k, rhs_p, u_p):
        try:
            B = np.ctypeslib.as_array(rhs_p, shape=(k, n)).copy()
            U = np.asarray(spsd_solve(B), dtype=np.float64).reshape(k, n)
            np.ctypeslib.as_array(u_p, shape=(k, n))[:] = U
            return 0
        except Exception as e:                            # reported through the C++ exception path
            failure.append(e)
            return 1
    cb = CB(_cb)
    dp = POINTER(c_double)
    lib.mfemhost_constrained_solve.argtypes = [c_int64, c_int, dp, dp, c_int64, POINTER(c_int64), c_int, dp, c_int, dp, dp, dp, CB]
    lib.mfemhost_constrained_solve.restype = c_int
    rc = lib.mfemhost_constrained_solve(n, m, rows.ctypes.data_as(dp), rows_rhs.ctypes.data_as(dp), fixed_vars.size,
                                        fixed_vars.ctypes.data_as(POINTER(c_int64)), rigid_modes.shape[0] if rigid_modes.size else 0,
                                        rigid_modes.ctypes.data_as(dp), nrhs, fs.ctypes.data_as(dp), us.ctypes.data_as(dp),
                                        lam.ctypes.data_as(dp), cb)
    if rc != 0:
        if failure:
            raise failure[0]
        raise _err(lib)
    return us, lam


def m2m_tensor(dim, E, G, S):
    """E : G : S for flattened rank-4 tensors (G without major symmetry) as PeriodicHomogenization_cli --m2mstress
    computes and prints it: (flattened result, Mathematica-array text)."""
    lib = load_library()
    F = dim * (dim + 1) // 2
    dp = POINTER(c_double)
    a = [np.ascontiguousarray(x, dtype=np.float64).reshape(F, F) for x in (E, G, S)]
    out = np.zeros((F, F))
    text = ctypes.create_string_buffer(1 << 14)
    lib.mfemhost_m2m_tensor.argtypes = [c_int, dp, dp, dp, dp, ctypes.c_char_p, c_int]
    if lib.mfemhost_m2m_tensor(dim, a[0].ctypes.data_as(dp), a[1].ctypes.data_as(dp), a[2].ctypes.data_as(dp), out.ctypes.data_as(dp),
                               text, len(text)) != 0:
        raise _err(lib)
    return out, text.value.decode()
