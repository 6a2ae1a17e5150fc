"""Python access to oracle/ref_cpu.cc (the C++ restatement of the reference's assembly loop
structure).  TEST INFRASTRUCTURE: only tests/, __graft_entry__ and bench.py's cpu_baseline /
--impl reference legs may import this."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_int, c_int32, c_int64, c_void_p

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "ref_cpu.cc")
_LIB = os.path.join(_HERE, "_build", "libref_cpu.so")
_lib = None


def build(force=False):
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-pthread", _SRC, "-o", _LIB])
    return _LIB


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.refcpu_assemble.restype = c_void_p
        lib.refcpu_assemble.argtypes = [c_int, c_int, c_int64, POINTER(c_double), c_int64, POINTER(c_int32),
                                        POINTER(c_double), c_int, POINTER(c_int64), c_int64, c_int,
                                        POINTER(c_double), POINTER(c_int64)]
        lib.refcpu_copy.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_double)]
        lib.refcpu_free.argtypes = [c_void_p]
        lib.refcpu_pcg.argtypes = [c_int64, POINTER(c_int64), POINTER(c_int32), POINTER(c_double), c_int,
                                   POINTER(c_double), POINTER(c_double), POINTER(c_double), c_double, c_int, c_int,
                                   POINTER(c_int), POINTER(c_double), POINTER(c_double)]
        lib.refcpu_element_stiffness.argtypes = [c_int, c_int, POINTER(c_double), POINTER(c_double), POINTER(c_double)]
        _lib = lib
    return _lib


def assemble_upper_csc(N, deg, nodes, elem_nodes, D, dof_for_node=None, n_dofs=None, threads=None):
    """Upper-triangle CSC of K exactly as m_assembleStiffnessMatrix + sumRepeated produce it.
    Returns (csc_matrix, dict(ke=..., scatter=..., compress=...))."""
    lib = _load()
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    en = np.ascontiguousarray(elem_nodes, dtype=np.int32)
    D = np.ascontiguousarray(D, dtype=np.float64)
    per = 1 if D.ndim == 3 else 0
    threads = threads or os.cpu_count() or 1
    dofp = None
    nd = nodes.shape[0]
    if dof_for_node is not None:
        dof = np.ascontiguousarray(dof_for_node, dtype=np.int64)
        dofp = dof.ctypes.data_as(POINTER(c_int64))
        nd = int(n_dofs)
    t = (c_double * 3)()
    nnz = c_int64()
    res = lib.refcpu_assemble(N, deg, nodes.shape[0], nodes.ctypes.data_as(POINTER(c_double)), en.shape[0],
                              en.ctypes.data_as(POINTER(c_int32)), D.ctypes.data_as(POINTER(c_double)), per, dofp, nd,
                              threads, t, ctypes.byref(nnz))
    if not res:
        raise RuntimeError("refcpu_assemble failed")
    n = N * nd
    colptr = np.zeros(n + 1, dtype=np.int64); rowidx = np.zeros(nnz.value, dtype=np.int64); vals = np.zeros(nnz.value)
    lib.refcpu_copy(res, colptr.ctypes.data_as(POINTER(c_int64)), rowidx.ctypes.data_as(POINTER(c_int64)),
                    vals.ctypes.data_as(POINTER(c_double)))
    lib.refcpu_free(res)
    A = sp.csc_matrix((vals, rowidx, colptr), shape=(n, n))
    return A, dict(ke=t[0], scatter=t[1], compress=t[2], threads=threads)


def element_stiffness(N, deg, pts, D):
    lib = _load()
    nn = (N + 1) if deg == 1 else (6 if N == 2 else 10)
    Ke = np.full((N * nn, N * nn), np.nan)
    pts = np.ascontiguousarray(pts, dtype=np.float64); D = np.ascontiguousarray(D, dtype=np.float64)
    lib.refcpu_element_stiffness(N, deg, pts.ctypes.data_as(POINTER(c_double)), D.ctypes.data_as(POINTER(c_double)),
                                 Ke.ctypes.data_as(POINTER(c_double)))
    return Ke


def solve_fixed_pcg(N, Aupper, f, fixed_vars, fixed_vals, rtol=1e-8, max_iters=200000, threads=None):
    """K_ff u_f = f_f - K_fc u_c by block-Jacobi PCG on the host cores (fixed variables become
    identity rows, exactly the GPU solver's formulation).  Returns (u, dict(iters, seconds, relres))."""
    lib = _load()
    threads = threads or os.cpu_count() or 1
    n = Aupper.shape[0]
    K = (Aupper + sp.triu(Aupper, 1).T).tocsr()
    fixed_vars = np.asarray(fixed_vars, dtype=np.int64)
    u = np.zeros(n); u[fixed_vars] = fixed_vals
    free = np.ones(n); free[fixed_vars] = 0.0
    b = (np.asarray(f, float).reshape(-1) - K @ u) * free
    Dm = sp.diags(free)
    Kmask = (Dm @ K @ Dm + sp.diags(1.0 - free)).tocsr()
    Kmask.sort_indices()
    nb = n // N
    bsr = Kmask.tobsr((N, N)); bsr.sort_indices()
    Minv = np.zeros((nb, N, N))
    for i in range(nb):
        cols = bsr.indices[bsr.indptr[i]:bsr.indptr[i + 1]]
        k = bsr.indptr[i] + np.searchsorted(cols, i)
        Minv[i] = np.linalg.inv(bsr.data[k])
    x = np.zeros(n)
    rowptr = Kmask.indptr.astype(np.int64); colidx = Kmask.indices.astype(np.int32); vals = Kmask.data
    iters = c_int(); secs = c_double(); relres = c_double()
    st = lib.refcpu_pcg(n, rowptr.ctypes.data_as(POINTER(c_int64)), colidx.ctypes.data_as(POINTER(c_int32)),
                        vals.ctypes.data_as(POINTER(c_double)), N, Minv.ctypes.data_as(POINTER(c_double)),
                        b.ctypes.data_as(POINTER(c_double)), x.ctypes.data_as(POINTER(c_double)), rtol, max_iters,
                        threads, ctypes.byref(iters), ctypes.byref(secs), ctypes.byref(relres))
    if st != 0:
        raise RuntimeError(f"refcpu_pcg failed with status {st}")
    return x + u, dict(iters=iters.value, seconds=secs.value, relres=relres.value, threads=threads)
