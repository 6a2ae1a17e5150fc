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
