"""ctypes binding of the C ABI in include/mfem_b200.h.

This is the same binding a MeshFEM maintainer would write for the Python module
(src/python_bindings/*.cc call the same entry points from C++).
There is no CPU fallback: if libmfem_b200.so is missing or no CUDA device is
usable, construction fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmfem_b200.so")

SYMBOLS = [
    "mfem_b200_create", "mfem_b200_destroy", "mfem_b200_last_error", "mfem_b200_device_count",
    "mfem_b200_comm_unique_id", "mfem_b200_comm_init", "mfem_b200_set_option", "mfem_b200_set_mesh",
    "mfem_b200_set_interface", "mfem_b200_set_material_constant", "mfem_b200_set_material_per_element",
    "mfem_b200_assemble", "mfem_b200_get_bsr_sizes", "mfem_b200_get_bsr", "mfem_b200_dump_upper_triplets",
    "mfem_b200_set_node_positions", "mfem_b200_fix_variables", "mfem_b200_clear_fixed_variables",
    "mfem_b200_solve", "mfem_b200_apply_K", "mfem_b200_spmv", "mfem_b200_const_strain_load",
    "mfem_b200_avg_strain_stress", "mfem_b200_get_volumes", "mfem_b200_get_timer", "mfem_b200_reset_timers",
    "mfem_b200_launch_count", "mfem_b200_time_spmv", "mfem_b200_time_operator", "mfem_b200_set_matrix_triplets", "mfem_b200_comm_share",
    "mfem_b200_apply_preconditioner", "mfem_b200_get_coarse_array", "mfem_b200_release_cached_memory",
    "mfem_b200_comm_window_handle", "mfem_b200_comm_window_open", "mfem_b200_comm_uses_peer_window",
    "mfem_b200_apply_delta_K", "mfem_b200_delta_const_strain_load", "mfem_b200_delta_avg_strain",
]

STATUS_NAMES = {
    0: "OK", -1: "ERR_INVALID", -2: "ERR_CUDA", -3: "ERR_NEG_VOLUME", -4: "ERR_ALREADY_FIXED",
    -5: "ERR_BAD_RHS", -6: "ERR_NOT_SPD", -7: "ERR_NO_CONVERGE", -8: "ERR_NAN", -9: "ERR_COMM",
}


class SolveInfo(ctypes.Structure):
    _fields_ = [("iterations", c_int32), ("converged", c_int32), ("rel_residual", c_double),
                ("seconds", c_double), ("spmv_seconds", c_double)]


class MfemB200Error(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"[{STATUS_NAMES.get(status, status)}] {msg}")
        self.status = status
        self.message = msg


_lib = None


def load_library():
    """Load libmfem_b200.so (building is the job of __graft_entry__.build / meshfem_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m meshfem_b200.build` "
            "(there is no CPU fallback for the assemble-and-solve path)")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    dp, ip, lp = POINTER(c_double), POINTER(c_int32), POINTER(c_int64)
    lib.mfem_b200_create.argtypes = [c_int, POINTER(c_void_p)]
    lib.mfem_b200_destroy.argtypes = [c_void_p]
    lib.mfem_b200_last_error.argtypes = [c_void_p]
    lib.mfem_b200_last_error.restype = c_char_p
    lib.mfem_b200_device_count.argtypes = []
    lib.mfem_b200_comm_unique_id.argtypes = [c_void_p]
    lib.mfem_b200_comm_init.argtypes = [c_void_p, c_int, c_int, c_void_p]
    lib.mfem_b200_comm_share.argtypes = [c_void_p, c_void_p]
    lib.mfem_b200_comm_window_handle.argtypes = [c_void_p, c_void_p]
    lib.mfem_b200_comm_window_open.argtypes = [c_void_p, c_void_p]
    lib.mfem_b200_comm_uses_peer_window.argtypes = [c_void_p]
    lib.mfem_b200_set_option.argtypes = [c_void_p, c_char_p, c_int64]
    lib.mfem_b200_set_mesh.argtypes = [c_void_p, c_int, c_int, c_int64, dp, c_int64, ip, lp, c_int64]
    lib.mfem_b200_set_interface.argtypes = [c_void_p, c_int, ip, lp, ip, POINTER(ctypes.c_uint8)]
    lib.mfem_b200_set_material_constant.argtypes = [c_void_p, dp]
    lib.mfem_b200_set_material_per_element.argtypes = [c_void_p, dp]
    lib.mfem_b200_assemble.argtypes = [c_void_p]
    lib.mfem_b200_get_bsr_sizes.argtypes = [c_void_p, lp, lp]
    lib.mfem_b200_get_bsr.argtypes = [c_void_p, lp, ip, dp]
    lib.mfem_b200_dump_upper_triplets.argtypes = [c_void_p, c_char_p]
    lib.mfem_b200_set_node_positions.argtypes = [c_void_p, dp]
    lib.mfem_b200_fix_variables.argtypes = [c_void_p, c_int64, lp, dp]
    lib.mfem_b200_clear_fixed_variables.argtypes = [c_void_p]
    lib.mfem_b200_solve.argtypes = [c_void_p, c_int, dp, dp, c_double, c_int, POINTER(SolveInfo)]
    lib.mfem_b200_apply_K.argtypes = [c_void_p, dp, dp]
    lib.mfem_b200_spmv.argtypes = [c_void_p, dp, dp]
    lib.mfem_b200_const_strain_load.argtypes = [c_void_p, dp, dp]
    lib.mfem_b200_avg_strain_stress.argtypes = [c_void_p, dp, dp, dp]
    lib.mfem_b200_get_volumes.argtypes = [c_void_p, dp]
    lib.mfem_b200_get_timer.argtypes = [c_void_p, c_char_p]
    lib.mfem_b200_get_timer.restype = c_double
    lib.mfem_b200_reset_timers.argtypes = [c_void_p]
    lib.mfem_b200_launch_count.argtypes = [c_void_p]
    lib.mfem_b200_launch_count.restype = c_int64
    lib.mfem_b200_time_spmv.argtypes = [c_void_p, c_int, dp]
    lib.mfem_b200_time_operator.argtypes = [c_void_p, c_int, dp, dp, ctypes.POINTER(c_int)]
    lib.mfem_b200_apply_preconditioner.argtypes = [c_void_p, dp, dp, dp]
    lib.mfem_b200_get_coarse_array.argtypes = [c_void_p, c_char_p, dp, c_int64, lp]
    lib.mfem_b200_apply_delta_K.argtypes = [c_void_p, dp, dp, c_int64, dp]
    lib.mfem_b200_delta_const_strain_load.argtypes = [c_void_p, dp, dp, c_int64, dp]
    lib.mfem_b200_delta_avg_strain.argtypes = [c_void_p, dp, dp, dp, c_int64, dp]
    lib.mfem_b200_set_matrix_triplets.argtypes = [c_void_p, c_int, c_int64, c_int64, lp, lp, dp, c_int]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is c_int and name not in ("mfem_b200_device_count",):
            fn.restype = c_int
    _lib = lib
    return lib


def _dptr(a):
    return a.ctypes.data_as(POINTER(c_double))


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def nodes_per_elem(dim, deg):
    return dim + 1 if deg == 1 else (6 if dim == 2 else 10)


def flat_len(dim):
    return dim * (dim + 1) // 2


class Handle:
    """Thin object wrapper over an mfem_b200_handle."""

    def __init__(self, device: int = 0, **options):
        self.lib = load_library()
        self._h = c_void_p()
        st = self.lib.mfem_b200_create(device, ctypes.byref(self._h))
        if st != 0:
            msg = self.lib.mfem_b200_last_error(None).decode()
            self._h = None
            raise MfemB200Error(st, msg)
        self.dim = self.deg = 0
        self.n_nodes = self.n_elems = self.n_dofs = 0
        for k, v in options.items():
            self.set_option(k, v)

    def _check(self, st):
        if st != 0:
            raise MfemB200Error(st, self.lib.mfem_b200_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.mfem_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- configuration -----------------------------------------------------
    def set_option(self, name, value):
        self._check(self.lib.mfem_b200_set_option(self._h, name.encode(), int(value)))

    def comm_init(self, n_ranks, rank, unique_id: bytes):
        buf = ctypes.create_string_buffer(unique_id, 128)
        self._check(self.lib.mfem_b200_comm_init(self._h, n_ranks, rank, buf))

    def comm_share(self, parent):
        """Use the communicator of `parent` (a long-lived handle of this process) instead of creating one."""
        self._check(self.lib.mfem_b200_comm_share(self._h, parent._h))
        self._comm_parent = parent          # keep it alive

    def comm_window_handle(self) -> bytes:
        """Allocate this rank's peer window and return its 64-byte CUDA IPC handle (after comm_init)."""
        buf = ctypes.create_string_buffer(64)
        self._check(self.lib.mfem_b200_comm_window_handle(self._h, buf))
        return buf.raw

    def comm_window_open(self, handles_rank_order: bytes):
        """Map the windows of all ranks (their IPC handles concatenated in rank order)."""
        buf = ctypes.create_string_buffer(handles_rank_order, len(handles_rank_order))
        self._check(self.lib.mfem_b200_comm_window_open(self._h, buf))

    def comm_uses_peer_window(self) -> bool:
        return bool(self.lib.mfem_b200_comm_uses_peer_window(self._h))

    @staticmethod
    def comm_unique_id() -> bytes:
        lib = load_library()
        buf = ctypes.create_string_buffer(128)
        st = lib.mfem_b200_comm_unique_id(buf)
        if st != 0:
            raise MfemB200Error(st, "ncclGetUniqueId failed")
        return buf.raw

    def set_mesh(self, dim, deg, nodes, elem_nodes, dof_for_node=None, n_dofs=None):
        nodes = _f64(nodes)
        en = np.ascontiguousarray(elem_nodes, dtype=np.int32)
        assert nodes.ndim == 2 and nodes.shape[1] == dim
        assert en.ndim == 2 and en.shape[1] == nodes_per_elem(dim, deg)
        dptr = None
        nd = nodes.shape[0]
        if dof_for_node is not None:
            dof = np.ascontiguousarray(dof_for_node, dtype=np.int64)
            dptr = dof.ctypes.data_as(POINTER(c_int64))
            nd = int(n_dofs if n_dofs is not None else dof.max() + 1)
        self._check(self.lib.mfem_b200_set_mesh(self._h, dim, deg, nodes.shape[0], _dptr(nodes), en.shape[0],
                                                en.ctypes.data_as(POINTER(c_int32)), dptr, nd))
        self.dim, self.deg = dim, deg
        self.n_nodes, self.n_elems, self.n_dofs = nodes.shape[0], en.shape[0], nd

    def set_interface(self, neighbor_ranks, neighbor_offsets, shared_local_dofs, owned):
        nr = np.ascontiguousarray(neighbor_ranks, dtype=np.int32)
        no = np.ascontiguousarray(neighbor_offsets, dtype=np.int64)
        sh = np.ascontiguousarray(shared_local_dofs, dtype=np.int32)
        ow = np.ascontiguousarray(owned, dtype=np.uint8)
        assert ow.size == self.n_dofs
        self._check(self.lib.mfem_b200_set_interface(self._h, nr.size, nr.ctypes.data_as(POINTER(c_int32)),
                                                     no.ctypes.data_as(POINTER(c_int64)),
                                                     sh.ctypes.data_as(POINTER(c_int32)),
                                                     ow.ctypes.data_as(POINTER(ctypes.c_uint8))))

    def set_node_positions(self, nodes):
        nodes = _f64(nodes, (self.n_nodes, self.dim))
        self._check(self.lib.mfem_b200_set_node_positions(self._h, _dptr(nodes)))

    def set_material(self, D):
        D = _f64(D)
        F = flat_len(self.dim)
        if D.shape == (F, F):
            self._check(self.lib.mfem_b200_set_material_constant(self._h, _dptr(D)))
        elif D.shape == (self.n_elems, F, F):
            self._check(self.lib.mfem_b200_set_material_per_element(self._h, _dptr(D)))
        else:
            raise ValueError(f"material tensor has shape {D.shape}")

    # ---- assembly ------------------------------------------------------------
    def assemble(self):
        self._check(self.lib.mfem_b200_assemble(self._h))

    def set_matrix(self, K, block_dim=3, upper_triangle_only=False):
        """SPSDSystem(K) for a matrix assembled elsewhere: K = scipy sparse matrix (any format; summed COO
        semantics) or a tuple (n_vars, rows, cols, vals)."""
        if isinstance(K, tuple):
            n, rows, cols, vals = K
        else:
            coo = K.tocoo()
            n, rows, cols, vals = coo.shape[0], coo.row, coo.col, coo.data
        rows = np.ascontiguousarray(rows, dtype=np.int64); cols = np.ascontiguousarray(cols, dtype=np.int64)
        vals = _f64(vals)
        self._check(self.lib.mfem_b200_set_matrix_triplets(self._h, block_dim, int(n), rows.size,
                                                           rows.ctypes.data_as(POINTER(c_int64)),
                                                           cols.ctypes.data_as(POINTER(c_int64)), _dptr(vals),
                                                           1 if upper_triangle_only else 0))
        self.dim, self.deg = block_dim, 0
        self.n_nodes = self.n_dofs = int(n) // block_dim
        self.n_elems = 0

    def bsr_sizes(self):
        nb, nnzb = c_int64(), c_int64()
        self._check(self.lib.mfem_b200_get_bsr_sizes(self._h, ctypes.byref(nb), ctypes.byref(nnzb)))
        return nb.value, nnzb.value

    def get_bsr(self, values=True):
        nb, nnzb = self.bsr_sizes()
        N = self.dim
        rowptr = np.zeros(nb + 1, dtype=np.int64)
        colidx = np.zeros(nnzb, dtype=np.int32)
        vals = np.zeros((nnzb, N, N)) if values else None
        self._check(self.lib.mfem_b200_get_bsr(self._h, rowptr.ctypes.data_as(POINTER(c_int64)),
                                               colidx.ctypes.data_as(POINTER(c_int32)),
                                               _dptr(vals) if values else None))
        return rowptr, colidx, vals

    def get_matrix(self):
        """Assembled K as a scipy BSR matrix in the caller's numbering."""
        import scipy.sparse as sp
        rowptr, colidx, vals = self.get_bsr()
        n = self.n_dofs * self.dim
        return sp.bsr_matrix((vals, colidx, rowptr), shape=(n, n))

    def dump_upper_triplets(self, path):
        self._check(self.lib.mfem_b200_dump_upper_triplets(self._h, os.fsencode(path)))

    # ---- constraints + solve ----------------------------------------------------
    def fix_variables(self, vars_, values=None):
        v = np.ascontiguousarray(vars_, dtype=np.int64)
        x = None if values is None else _f64(values)
        self._check(self.lib.mfem_b200_fix_variables(self._h, v.size, v.ctypes.data_as(POINTER(c_int64)),
                                                     None if x is None else _dptr(x)))

    def clear_fixed_variables(self):
        self._check(self.lib.mfem_b200_clear_fixed_variables(self._h))

    def solve(self, f, rtol=1e-10, max_iters=100000, return_info=False, out=None):
        """`out`: optional result array (same size as f, C-contiguous float64), e.g. page-locked memory reused across
        solves -- a fresh pageable array costs its page faults plus a staged device-to-host copy (~0.1 s per 350 MB)."""
        n = self.n_dofs * self.dim
        f = _f64(f)
        nrhs = f.size // n
        assert f.size == nrhs * n and nrhs >= 1
        if out is None:
            u = np.empty_like(f)
        else:
            assert isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous and out.size == f.size
            u = out
        info = (SolveInfo * nrhs)()
        st = self.lib.mfem_b200_solve(self._h, nrhs, _dptr(f), _dptr(u), rtol, max_iters, info)
        self.last_info = [dict(iterations=i.iterations, converged=bool(i.converged), rel_residual=i.rel_residual,
                               seconds=i.seconds) for i in info]
        self._check(st)
        return (u, self.last_info) if return_info else u

    # ---- operators -----------------------------------------------------------------
    def spmv(self, x):
        x = _f64(x)
        y = np.zeros_like(x)
        self._check(self.lib.mfem_b200_spmv(self._h, _dptr(x), _dptr(y)))
        return y

    def apply_K(self, u_nodes):
        u = _f64(u_nodes)
        out = np.zeros_like(u)
        self._check(self.lib.mfem_b200_apply_K(self._h, _dptr(u), _dptr(out)))
        return out

    def const_strain_load(self, eps_flat):
        e = _f64(eps_flat)
        out = np.zeros((self.n_dofs, self.dim))
        self._check(self.lib.mfem_b200_const_strain_load(self._h, _dptr(e), _dptr(out)))
        return out

    def avg_strain_stress(self, u_nodes):
        u = _f64(u_nodes)
        F = flat_len(self.dim)
        strain = np.zeros((self.n_elems, F))
        stress = np.zeros((self.n_elems, F))
        self._check(self.lib.mfem_b200_avg_strain_stress(self._h, _dptr(u), _dptr(strain), _dptr(stress)))
        return strain, stress

    # ---- discrete shape derivatives (per-vertex perturbation delta_p[n_vertices, dim]) ---------------------
    def apply_delta_K(self, u_nodes, delta_p):
        u, dpv = _f64(u_nodes), _f64(delta_p)
        out = np.zeros((self.n_dofs, self.dim))
        self._check(self.lib.mfem_b200_apply_delta_K(self._h, _dptr(u), _dptr(dpv), dpv.size // self.dim, _dptr(out)))
        return out

    def delta_const_strain_load(self, eps_flat, delta_p):
        e, dpv = _f64(eps_flat), _f64(delta_p)
        out = np.zeros((self.n_dofs, self.dim))
        self._check(self.lib.mfem_b200_delta_const_strain_load(self._h, _dptr(e), _dptr(dpv), dpv.size // self.dim, _dptr(out)))
        return out

    def delta_avg_strain(self, u_nodes, delta_u_nodes, delta_p):
        u, du, dpv = _f64(u_nodes), _f64(delta_u_nodes), _f64(delta_p)
        out = np.zeros((self.n_elems, flat_len(self.dim)))
        self._check(self.lib.mfem_b200_delta_avg_strain(self._h, _dptr(u), _dptr(du), _dptr(dpv), dpv.size // self.dim, _dptr(out)))
        return out

    def volumes(self):
        v = np.zeros(self.n_elems)
        self._check(self.lib.mfem_b200_get_volumes(self._h, _dptr(v)))
        return v

    # ---- bookkeeping ----------------------------------------------------------------
    def timer(self, section):
        return self.lib.mfem_b200_get_timer(self._h, section.encode())

    def reset_timers(self):
        self.lib.mfem_b200_reset_timers(self._h)

    def launch_count(self):
        return self.lib.mfem_b200_launch_count(self._h)

    def apply_preconditioner(self, r):
        """(M^-1 r, r.M^-1 r) for the preconditioner the next solve would use (diagnostics / parity tests)."""
        r = _f64(r)
        z = np.zeros_like(r)
        rz = c_double()
        self._check(self.lib.mfem_b200_apply_preconditioner(self._h, _dptr(r), _dptr(z), ctypes.byref(rz)))
        return z, rz.value

    def coarse_array(self, name):
        """Named array of the aggregation levels (diagnostics; see mfem_b200_get_coarse_array)."""
        n = c_int64()
        self._check(self.lib.mfem_b200_get_coarse_array(self._h, name.encode(), None, 0, ctypes.byref(n)))
        out = np.zeros(n.value)
        self._check(self.lib.mfem_b200_get_coarse_array(self._h, name.encode(), _dptr(out), n.value, ctypes.byref(n)))
        return out

    def time_spmv(self, iters=20):
        s = c_double()
        self._check(self.lib.mfem_b200_time_spmv(self._h, iters, ctypes.byref(s)))
        return s.value

    def time_operator(self, iters=20):
        """Seconds per K*p product as the PCG launches it: (seconds, [element kernel, gather kernel], matrix_free)."""
        s = c_double()
        parts = (c_double * 2)()
        mf = c_int(0)
        self._check(self.lib.mfem_b200_time_operator(self._h, iters, ctypes.byref(s), parts, ctypes.byref(mf)))
        return s.value, [parts[0], parts[1]], bool(mf.value)
