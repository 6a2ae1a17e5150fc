"""CPU emulation (numpy/scipy) of exactly what csrc/coarse.inl does on one GPU: bounding box of the DoF positions, a
near-cubic LARGE box grid within the budget (coarse_choose_boxes) refined into SMALL boxes of about `fine` nodes
(coarse_choose_refinement), level-1 slots (large-box-major), positions centred per small box, the shift table
d = centre(small) - centre(large), rigid-body modes masked on the fixed variables,
    M^-1 = B0^-1 + P1 B1^-1 P1' + Z2 E2^-1 Z2',   Z2 = P1 P2,
E2 = Z2'K_ff Z2 with a unit diagonal on dead modes and a 1e-8 relative shift (explicit inverse), B1 = pseudo-inverses
of the diagonal blocks of P1'K_ff P1 (eigenvalues below 1e-10 of the largest dropped), all inside PCG with
r.z = r.B0^-1 r + c1.y1 + c2.y2.  The level-1 <-> level-2 transfers use the device formulas (coarse_shift_restrict /
coarse_shift_prolong), and |P1 P2 - Z2| is checked.  Prints block-Jacobi, two-level (fine = 0) and multilevel counts;
tests/test_coarse_logic.py runs a small case.

  python tools/emulate_multilevel.py            # the three cases of tests/test_zz_two_level_gpu.py
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


def choose_refinement(L, b, ndofs, target):
    L = np.asarray(L, float); nz = L > 0
    r = [1] * len(L)
    if target <= 0 or not nz.any() or ndofs <= 0:
        return r
    h1 = (np.prod(L[nz]) * target / ndofs) ** (1.0 / nz.sum())
    for k in range(len(L)):
        if L[k] > 0:
            r[k] = int(max(1, np.floor(L[k] / b[k] / h1 + 0.5)))
    return r


def rigid(Y):
    nd, N = Y.shape
    M = 6 if N == 3 else 3
    R = np.zeros((nd, N, M))
    for k in range(N): R[:, k, k] = 1.0
    if N == 3:
        R[:, 1, 3], R[:, 2, 3] = -Y[:, 2], Y[:, 1]
        R[:, 0, 4], R[:, 2, 4] = Y[:, 2], -Y[:, 0]
        R[:, 0, 5], R[:, 1, 5] = -Y[:, 1], Y[:, 0]
    else:
        R[:, 0, 2], R[:, 1, 2] = -Y[:, 1], Y[:, 0]
    return R


def shift_restrict(d, c):                 # coarse_shift_restrict: rows of c are slot coefficient vectors
    c = c.copy()
    if d.shape[1] == 3:
        c[:, 3] += d[:, 1] * c[:, 2] - d[:, 2] * c[:, 1]
        c[:, 4] += d[:, 2] * c[:, 0] - d[:, 0] * c[:, 2]
        c[:, 5] += d[:, 0] * c[:, 1] - d[:, 1] * c[:, 0]
    else:
        c[:, 2] += d[:, 0] * c[:, 1] - d[:, 1] * c[:, 0]
    return c


def shift_prolong(d, Y):                  # coarse_shift_prolong
    q = np.zeros_like(Y)
    if d.shape[1] == 3:
        q[:, 0] = Y[:, 0] + Y[:, 4] * d[:, 2] - Y[:, 5] * d[:, 1]
        q[:, 1] = Y[:, 1] + Y[:, 5] * d[:, 0] - Y[:, 3] * d[:, 2]
        q[:, 2] = Y[:, 2] + Y[:, 3] * d[:, 1] - Y[:, 4] * d[:, 0]
        q[:, 3:] = Y[:, 3:]
    else:
        q[:, 0] = Y[:, 0] - Y[:, 2] * d[:, 1]
        q[:, 1] = Y[:, 1] + Y[:, 2] * d[:, 0]
        q[:, 2] = Y[:, 2]
    return q


def dropping_cholesky_inverse(A, rel_tol=1e-10):
    """k_coarse_invert1: inverse of the principal submatrix of the modes whose Cholesky pivot exceeds rel_tol * max diag,
    embedded in zeros."""
    M = A.shape[0]
    L = np.array(A, float)
    dmax = max(np.diag(L).max(), 0.0)
    tol = rel_tol * dmax
    for k in range(M):
        d = L[k, k] - L[k, :k] @ L[k, :k]
        live = d > tol and dmax > 0.0
        piv = np.sqrt(d) if live else 0.0
        L[k, k] = piv
        for i in range(k + 1, M):
            L[i, k] = (L[i, k] - L[i, :k] @ L[k, :k]) / piv if live else 0.0
    W = np.zeros((M, M))
    for c in range(M):
        if L[c, c] <= 0.0:
            continue
        W[c, c] = 1.0 / L[c, c]
        for i in range(c + 1, M):
            W[i, c] = -(L[i, c:i] @ W[c:i, c]) / L[i, i] if L[i, i] > 0.0 else 0.0
    return W.T @ W


def build(N, X, free, Km, Sopt, fine):
    """Returns apply(r) -> (z_coarse, rz_coarse) and a dict of diagnostics."""
    nd = X.shape[0]; n = nd * N
    M = 6 if N == 3 else 3
    Sr = min(min(Sopt, 32768 // M), max(1, nd // 8))
    lo, hi = X.min(0), X.max(0)
    L = hi - lo
    b = np.array(choose_boxes(L, Sr))
    r = np.array(choose_refinement(L, b, nd, fine))
    scale1 = np.where(L > 0, b * r / np.where(L > 0, L, 1), 0.0)
    q = np.clip(np.floor((X - lo) * scale1).astype(np.int64), 0, b * r - 1)
    big = np.zeros(nd, dtype=np.int64); loc = np.zeros(nd, dtype=np.int64)
    for k in range(N):
        big = big * b[k] + q[:, k] // r[k]
        loc = loc * r[k] + q[:, k] % r[k]
    S2 = int(np.prod(b)); R = int(np.prod(r)); level1 = fine > 0 and R > 1
    S1 = S2 * R if level1 else 0
    cen2 = np.zeros((S2, N)); np.add.at(cen2, big, X); cnt2 = np.bincount(big, minlength=S2); cen2 /= np.maximum(cnt2, 1)[:, None]
    Y2 = X - cen2[big]
    if level1:
        slot = big * R + loc
        cen1 = np.zeros((S1, N)); np.add.at(cen1, slot, X); cnt1 = np.bincount(slot, minlength=S1); cen1 /= np.maximum(cnt1, 1)[:, None]
        shift = np.where(((cnt1 > 0) & (cnt2[np.arange(S1) // R] > 0))[:, None], cen1 - cen2[np.arange(S1) // R], 0.0)
        Y1 = Y2 - shift[slot]
    else:
        slot = S1 + big; Y1 = Y2; shift = np.zeros((0, N))
    fm = np.repeat(free, M)
    rows = np.repeat(np.arange(n), M)
    Z2 = sp.csr_matrix((rigid(Y2).reshape(-1) * fm, (rows, (M * np.repeat(big, N)[:, None] + np.arange(M)[None, :]).reshape(-1))), shape=(n, M * S2))
    E = (Z2.T @ Km @ Z2).toarray(); d = np.diag(E).copy()
    E[np.diag_indices_from(E)] = np.where(d == 0.0, 1.0, d * (1 + 1e-8))
    Einv = sla.cho_solve(sla.cho_factor(E), np.eye(E.shape[0]))
    info = dict(b=b.tolist(), r=r.tolist(), S1=S1, S2=S2)
    arrays = dict(slot=slot, Y1=Y1, shift=shift, Einv=Einv)
    info["arrays"] = arrays
    if not level1:
        def apply(rv):
            c2 = Z2.T @ rv; y2 = Einv @ c2
            return Z2 @ y2, float(c2 @ y2)
        return apply, info
    P1 = sp.csr_matrix((rigid(Y1).reshape(-1) * fm, (rows, (M * np.repeat(slot, N)[:, None] + np.arange(M)[None, :]).reshape(-1))), shape=(n, M * S1))
    K1 = (P1.T @ Km @ P1).tobsr((M, M)); K1.sort_indices()
    B1inv = np.zeros((S1, M, M)); D1 = np.zeros((S1, M, M)); arrays["D1"] = D1
    for s in range(S1):
        cols = K1.indices[K1.indptr[s]:K1.indptr[s + 1]]
        k = np.searchsorted(cols, s)
        if k < cols.size and cols[k] == s:
            blk = K1.data[K1.indptr[s] + k]; blk = 0.5 * (blk + blk.T)
            D1[s] = blk
            B1inv[s] = dropping_cholesky_inverse(blk)
    arrays["B1inv"] = B1inv
    par = np.arange(S1) // R
    # the composite prolongation is the large boxes' rigid modes: P1 P2 = Z2
    blocks = np.zeros((S1, M, M))
    for m in range(M):
        e = np.zeros((S1, M)); e[:, m] = 1.0
        blocks[:, :, m] = shift_prolong(shift, e)
    P2 = sp.csr_matrix((blocks.reshape(-1), (np.repeat(np.arange(M * S1), M), (M * np.repeat(par, M)[:, None] + np.arange(M)[None, :]).reshape(-1))),
                       shape=(M * S1, M * S2))
    info["P1P2_minus_Z2"] = float(abs(P1 @ P2 - Z2).max())

    def apply(rv):
        c1 = (P1.T @ rv).reshape(S1, M)
        y1 = np.einsum("sab,sb->sa", B1inv, c1)
        rz1 = float((c1 * y1).sum())
        c2 = np.zeros((S2, M)); np.add.at(c2, par, shift_restrict(shift, c1)); c2 = c2.reshape(-1)
        y2 = Einv @ c2
        qv = y1 + shift_prolong(shift, y2.reshape(S2, M)[par])
        return P1 @ qv.reshape(-1), rz1 + float(c2 @ y2)
    return apply, info


def masked_system(sim, fixed):
    """(N, X, free, Km, Kff, jac): the masked matrix of the device PCG and its block-Jacobi."""
    N = sim.mesh.N
    K = sim.stiffness().tocsr(); n = K.shape[0]
    free = np.ones(n, bool); free[fixed] = False
    mask = sp.diags(free.astype(float))
    Km = (mask @ K @ mask).tocsr(); Kff = (Km + sp.diags((~free).astype(float))).tocsr()
    bs = Kff.tobsr((N, N)); bs.sort_indices(); nd = n // N
    Minv = np.zeros((nd, N, N))
    for i in range(nd):
        cols = bs.indices[bs.indptr[i]:bs.indptr[i + 1]]
        Minv[i] = np.linalg.inv(bs.data[bs.indptr[i] + np.searchsorted(cols, i)])
    jac = lambda r: np.einsum("bij,bj->bi", Minv, r.reshape(-1, N)).reshape(-1)
    return N, sim.mesh.nodes.reshape(nd, N), free, Km, Kff, jac


def full_operator(sim, fixed, Sopt, fine):
    """r -> (M^-1 r, r.M^-1 r) as mfem_b200_apply_preconditioner returns them (r masked on the fixed variables first)."""
    N, X, free, Km, Kff, jac = masked_system(sim, fixed)
    apply, info = build(N, X, free, Km, Sopt, fine) if Sopt else (lambda r: (0.0 * r, 0.0), {})

    def op(r):
        r = np.asarray(r, float).reshape(-1) * free
        zc, rzc = apply(r)
        zb = jac(r)
        return zb + zc, float(r @ zb) + rzc
    return op, info


def run(N, deg, sizes, Sopt, fine, rtol=1e-10, verbose=True):
    sim, fixed, vals, f = cantilever_problem(N, deg, sizes)
    N, X, free, Km, Kff, jac = masked_system(sim, fixed)
    b = f.reshape(-1) * free
    _, it0 = pcg(Kff, b, jac, rtol, 20000)
    out = [it0]
    for fn in ([0, fine] if fine else [0]):
        apply, info = build(N, X, free, Km, Sopt, fn)
        _, it = pcg(Kff, b, lambda r: jac(r) + apply(r)[0], rtol, 20000)
        out.append(it)
        if verbose:
            info.pop("arrays", None)
            print(N, deg, sizes, "budget", Sopt, "fine", fn, info, "jacobi", it0, "multilevel", it, "ratio %.2f" % (it / it0), flush=True)
    return out


if __name__ == "__main__":
    run(3, 2, (20, 4, 4), 128, 24)
    run(3, 1, (24, 6, 6), 64, 24)
    run(2, 2, (40, 8), 96, 24)
