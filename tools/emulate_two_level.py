"""CPU emulation (numpy/scipy) of exactly what csrc/coarse.inl does on one GPU with box aggregates: bounding box of the
DoF positions, near-cubic box grid within the budget (coarse_choose_boxes), box id per DoF (k_coarse_box_agg), positions
centred per box, rigid-body modes masked on the fixed variables, E = Z'K_ff Z with a unit diagonal on dead modes and a
1e-8 relative shift, explicit inverse, additive correction inside PCG with r.z += c.y.  Prints block-Jacobi and
two-level iteration counts; tests/test_coarse_logic.py runs a small case.

  python tools/emulate_two_level.py            # the three cases of tests/test_zz_two_level_gpu.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

from util import cantilever_problem
from proto_two_level import pcg

def choose_boxes(L, budget):
    L = np.asarray(L, float); nz = L > 0
    h = (np.prod(L[nz]) / max(budget, 1)) ** (1.0 / nz.sum()) if nz.any() else 1.0
    b = [int(max(1, np.floor(l / h + 0.5))) if l > 0 else 1 for l in L]     # llround
    while np.prod(b) > budget:
        k = int(np.argmax(b))
        if b[k] == 1: break
        b[k] -= 1
    return b

def run(N, deg, sizes, Sopt, rtol=1e-10, verbose=True):
    sim, fixed, vals, f = cantilever_problem(N, deg, sizes)
    K = sim.stiffness().tocsr(); n = K.shape[0]
    free = np.ones(n, bool); free[fixed] = False
    mask = sp.diags(free.astype(float))
    Km = (mask @ K @ mask).tocsr(); Kff = (Km + sp.diags((~free).astype(float))).tocsr()
    b = f.reshape(-1) * free
    bs = Kff.tobsr((N, N)); bs.sort_indices(); nd = n // N
    Minv = np.zeros((nd, N, N))
    for i in range(nd):
        cols = bs.indices[bs.indptr[i]:bs.indptr[i + 1]]
        Minv[i] = np.linalg.inv(bs.data[bs.indptr[i] + np.searchsorted(cols, i)])
    jac = lambda r: np.einsum("bij,bj->bi", Minv, r.reshape(-1, N)).reshape(-1)
    _, it0 = pcg(Kff, b, jac, rtol, 20000)
    X = sim.mesh.nodes.reshape(nd, N)
    M = 6 if N == 3 else 3
    Sr = min(min(Sopt, 32768 // M), max(1, nd // 8))
    lo, hi = X.min(0), X.max(0)
    bx = choose_boxes(hi - lo, Sr)
    scale = np.where(hi > lo, np.array(bx) / np.where(hi > lo, hi - lo, 1), 0.0)
    q = np.clip(np.floor((X - lo) * scale).astype(np.int64), 0, np.array(bx) - 1)
    agg = np.zeros(nd, dtype=np.int64)
    for k in range(N): agg = agg * bx[k] + q[:, k]
    S = Sr
    cen = np.zeros((S, N)); np.add.at(cen, agg, X); cnt = np.bincount(agg, minlength=S); cen /= np.maximum(cnt, 1)[:, None]
    Y = X - cen[agg]
    R = np.zeros((nd, N, M))
    for k in range(N): R[:, k, k] = 1.0
    if N == 3:
        R[:, 1, 3], R[:, 2, 3] = -Y[:, 2], Y[:, 1]
        R[:, 0, 4], R[:, 2, 4] = Y[:, 2], -Y[:, 0]
        R[:, 0, 5], R[:, 1, 5] = -Y[:, 1], Y[:, 0]
    else:
        R[:, 0, 2], R[:, 1, 2] = -Y[:, 1], Y[:, 0]
    rows = np.repeat(np.arange(n), M); cols = (M * np.repeat(agg, N)[:, None] + np.arange(M)[None, :]).reshape(-1)
    Z = sp.csr_matrix((R.reshape(-1) * np.repeat(free, M), (rows, cols)), shape=(n, M * S))
    E = (Z.T @ Km @ Z).toarray(); d = np.diag(E).copy()
    E[np.diag_indices_from(E)] = np.where(d == 0.0, 1.0, d * (1 + 1e-8))
    w = np.linalg.eigvalsh(E)
    Einv = sla.cho_solve(sla.cho_factor(E), np.eye(E.shape[0]))
    _, it1 = pcg(Kff, b, lambda r: jac(r) + Z @ (Einv @ (Z.T @ r)), rtol, 20000)
    if verbose:
        print(N, deg, sizes, "budget", Sopt, "boxes", bx, "used", int((cnt > 0).sum()), "jacobi", it0, "two-level", it1, "ratio %.2f" % (it1 / it0), "E eig min %.2e" % w[0], flush=True)
    return it0, it1



if __name__ == "__main__":
    run(3, 2, (20, 4, 4), 128)
    run(3, 1, (24, 6, 6), 64)
    run(2, 2, (40, 8), 96)
