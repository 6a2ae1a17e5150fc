"""CPU prototype (numpy/scipy; no GPU): additive THREE-level preconditioner for the cantilever workloads.

  M^-1 = B0^-1 + P1 B1^-1 P1' + P1 P2 E2^-1 P2' P1'

B0 = the 3x3 block-Jacobi of the device PCG; level 1 = rigid-body modes of SMALL aggregates (tens of nodes, contiguous
runs of the Morton order) with K1 = P1' K P1 (sparse, 6x6 blocks) and B1 = its 6x6 block-Jacobi; level 2 = rigid-body
modes of LARGE aggregates (groups of consecutive small ones) expressed in level-1 coordinates, E2 = P2' K1 P2 dense
and inverted exactly (what csrc/coarse.inl does today with one coarse level).  Everything is additive, so an
iteration still costs ONE fine SpMV; the extra work is segmented reductions / broadcasts over index ranges and 6x6
block solves.  Prints PCG iteration counts next to block-Jacobi and the two-level method of tools/proto_two_level.py.

  python tools/proto_three_level.py 40x8x8 2 30 64 [boxes]   # grid, degree, nodes per small aggregate, large aggregates;
                                                             # "boxes": nested near-cubic box grids instead of Morton runs
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

from proto_two_level import morton_order, pcg
from util import cantilever_problem


def rigid_modes(Y):
    """[n, 3, 6]: translations and infinitesimal rotations about the origin of Y."""
    R = np.zeros((Y.shape[0], 3, 6))
    R[:, 0, 0] = R[:, 1, 1] = R[:, 2, 2] = 1.0
    R[:, 1, 3], R[:, 2, 3] = -Y[:, 2], Y[:, 1]
    R[:, 0, 4], R[:, 2, 4] = Y[:, 2], -Y[:, 0]
    R[:, 0, 5], R[:, 1, 5] = -Y[:, 1], Y[:, 0]
    return R


def choose_boxes(L, budget):
    """Near-cubic box grid with at most `budget` boxes (csrc/coarse.inl coarse_choose_boxes)."""
    L = np.asarray(L, float)
    h = (np.prod(L) / max(budget, 1)) ** (1.0 / len(L))
    b = [int(max(1, np.floor(l / h + 0.5))) for l in L]
    while np.prod(b) > budget:
        k = int(np.argmax(b))
        if b[k] == 1:
            break
        b[k] -= 1
    return np.array(b)


def run(sizes, deg, nodes_per_small, S2, rtol=1e-8, boxes=False):
    sim, fixed, vals, f = cantilever_problem(3, deg, sizes)
    K = sim.stiffness().tocsr()
    n = K.shape[0]; N = 3; nd = n // N
    free = np.ones(n, bool); free[fixed] = False
    mask = sp.diags(free.astype(float))
    Km = (mask @ K @ mask).tocsr()
    Kff = (Km + sp.diags((~free).astype(float))).tocsr()
    b = f.reshape(-1) * free
    bs = Kff.tobsr((N, N)); bs.sort_indices()
    Minv = np.zeros((nd, N, N))
    for i in range(nd):
        cols = bs.indices[bs.indptr[i]:bs.indptr[i + 1]]
        Minv[i] = np.linalg.inv(bs.data[bs.indptr[i] + np.searchsorted(cols, i)])
    jac = lambda r: np.einsum("bij,bj->bi", Minv, r.reshape(-1, N)).reshape(-1)
    t = time.time()
    _, it0 = pcg(Kff, b, jac, rtol, 20000)
    print(f"grid {sizes} deg {deg}: {nd} nodes, block-Jacobi {it0} iterations ({time.time() - t:.1f}s)", flush=True)

    X = sim.mesh.nodes
    order = morton_order(X)
    pos = np.empty(nd, dtype=np.int64); pos[order] = np.arange(nd)

    lo, hi = X.min(0), X.max(0)
    if boxes:      # nested box grids: large boxes b2, small boxes bg1 = bg2 * r (so every small box lies in one large box)
        bg2 = choose_boxes(hi - lo, S2)
        r = max(1, int(round((nd / np.prod(bg2) / nodes_per_small) ** (1.0 / 3.0))))
        bg1 = bg2 * r
        S2 = int(np.prod(bg2))

    def box_ids(b):
        q = np.minimum(np.floor((X - lo) / (hi - lo) * b).astype(np.int64), b - 1)
        return q, (q[:, 0] * b[1] + q[:, 1]) * b[2] + q[:, 2]

    def level(S, b=None):
        agg = (pos * S) // nd if b is None else box_ids(b)[1]
        cen = np.zeros((S, 3)); np.add.at(cen, agg, X); cen /= np.maximum(np.bincount(agg, minlength=S), 1)[:, None]
        R = rigid_modes(X - cen[agg])
        rows = np.repeat(np.arange(n), 6)
        cols = (6 * np.repeat(agg, N)[:, None] + np.arange(6)[None, :]).reshape(-1)
        Z = sp.csr_matrix((R.reshape(-1) * np.repeat(free, 6), (rows, cols)), shape=(n, 6 * S))
        return agg, cen, Z

    def dense_inverse(E):
        d = np.diag(E).copy()
        E = E.copy(); E[np.diag_indices_from(E)] = np.where(d == 0.0, 1.0, d * (1 + 1e-8))
        return sla.cho_solve(sla.cho_factor(E), np.eye(E.shape[0]))

    # two-level with the large aggregates only (today's csrc/coarse.inl)
    _, _, Z2 = level(S2, bg2 if boxes else None)
    E2inv = dense_inverse((Z2.T @ Km @ Z2).toarray())
    _, it2 = pcg(Kff, b, lambda r: jac(r) + Z2 @ (E2inv @ (Z2.T @ r)), rtol, 20000)
    print(f"   two-level, {S2} aggregates ({nd / S2:.0f} nodes each): {it2} iterations", flush=True)

    # three-level: small aggregates (S1 a multiple of S2 so that the large ones are unions of consecutive small ones)
    S1 = int(np.prod(bg1)) if boxes else max(S2, int(round(nd / nodes_per_small / S2)) * S2)
    agg1, cen1, Z1 = level(S1, bg1 if boxes else None)
    K1 = (Z1.T @ Km @ Z1).tocsr()
    b1 = K1.tobsr((6, 6)); b1.sort_indices()
    B1inv = np.zeros((S1, 6, 6))
    for a in range(S1):
        cols = b1.indices[b1.indptr[a]:b1.indptr[a + 1]]
        blk = b1.data[b1.indptr[a] + np.searchsorted(cols, a)].copy()
        d = np.diag(blk).copy()
        blk[np.diag_indices(6)] = np.where(d == 0.0, 1.0, d * (1 + 1e-8))
        B1inv[a] = np.linalg.inv(blk)
    # level 2 in level-1 coordinates: small aggregate a in large aggregate A: t_a = T + W x (x_a - X_A), w_a = W
    if boxes:      # small box (i, j, k) of grid b1 lies in large box (i // r, j // r, k // r) of grid b2
        ii, jj, kk = np.meshgrid(np.arange(bg1[0]), np.arange(bg1[1]), np.arange(bg1[2]), indexing="ij")
        grp = (((ii // r) * bg2[1] + jj // r) * bg2[2] + kk // r).reshape(-1)
    else:
        grp = (np.arange(S1) * S2) // S1
    cen2 = np.zeros((S2, 3)); cnt1 = np.bincount(agg1, minlength=S1).astype(float)
    np.add.at(cen2, grp, np.nan_to_num(cen1) * cnt1[:, None]); cen2 /= np.maximum(np.bincount(grp, weights=cnt1, minlength=S2), 1)[:, None]
    dvec = cen1 - cen2[grp]
    blocks = np.zeros((S1, 6, 6))
    blocks[:, :3, :3] = np.eye(3); blocks[:, 3:, 3:] = np.eye(3)
    blocks[:, :3, 3:] = rigid_modes(dvec)[:, :, 3:]
    rows = np.repeat(np.arange(6 * S1), 6)
    cols = (6 * np.repeat(grp, 6)[:, None] + np.arange(6)[None, :]).reshape(-1)
    P2 = sp.csr_matrix((blocks.reshape(-1), (rows, cols)), shape=(6 * S1, 6 * S2))
    E3inv = dense_inverse((P2.T @ K1 @ P2).toarray())
    err = abs((Z1 @ P2) - Z2).max()      # the composite prolongation is the large aggregates' rigid modes

    def three(r):
        c1 = Z1.T @ r
        y1 = np.einsum("bij,bj->bi", B1inv, c1.reshape(-1, 6)).reshape(-1) + P2 @ (E3inv @ (P2.T @ c1))
        return jac(r) + Z1 @ y1

    _, it3 = pcg(Kff, b, three, rtol, 20000)
    nnzb1 = b1.indices.size
    print(f"   three-level, {S1} small aggregates ({nd / S1:.0f} nodes each, K1: {nnzb1} 6x6 blocks = "
          f"{nnzb1 * 36 / (bs.indices.size * 9):.3f} of K's values) + {S2} large: {it3} iterations "
          f"({it0 / it3:.1f}x vs block-Jacobi, {it2 / it3:.1f}x vs two-level; |Z1 P2 - Z2| = {err:.1e})", flush=True)
    # level 1 solved approximately by a FIXED Chebyshev polynomial in (M1^-1 K1), M1^-1 = B1^-1 + P2 E2^-1 P2'
    # (a fixed polynomial keeps the preconditioner a constant SPD operator, so plain PCG stays valid);
    # cost per outer iteration: 1 fine SpMV + k SpMVs with K1
    M1 = lambda c: np.einsum("bij,bj->bi", B1inv, c.reshape(-1, 6)).reshape(-1) + P2 @ (E3inv @ (P2.T @ c))
    # dead modes (fixed aggregates): K1 row is zero there; keep them out of the iteration
    live = (np.abs(K1).sum(axis=1).A1 > 0).astype(float)
    rng = np.random.default_rng(0)
    v = rng.standard_normal(6 * S1) * live
    for _ in range(30):
        w = M1(K1 @ v) * live; lmax = np.linalg.norm(w) / np.linalg.norm(v); v = w / np.linalg.norm(w)
    lmax *= 1.1
    for k, ratio in ((3, 10.0), (5, 20.0), (8, 40.0)):
        lmin = lmax / ratio
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)

        def cheb(c, k=k, theta=theta, delta=delta):
            # k steps of the Chebyshev iteration for K1 y = c from y = 0 (Saad, Alg. 12.1), preconditioned with M1
            c = c * live
            y = np.zeros_like(c); r = c.copy()
            sigma = theta / delta; rho = 1.0 / sigma
            d = M1(r) * live / theta
            for i in range(k):
                y = y + d
                r = r - K1 @ d
                rho_new = 1.0 / (2.0 * sigma - rho)
                d = rho_new * rho * d + (2.0 * rho_new / delta) * (M1(r) * live)
                rho = rho_new
            return y

        _, itc = pcg(Kff, b, lambda r: jac(r) + Z1 @ cheb(Z1.T @ r), rtol, 20000)
        cost = 1.0 + k * nnzb1 * 36 / (bs.indices.size * 9)
        print(f"   level 1 by Chebyshev degree {k} (lmax {lmax:.2f}, ratio {ratio:.0f}): {itc} outer iterations x {cost:.2f} fine-SpMV "
              f"equivalents = {itc * cost:.0f}  (block-Jacobi {it0}, two-level {it2}, additive three-level {it3})", flush=True)
    # level 1-2 V-cycle as the level-1 solver: k damped block-Jacobi (Chebyshev in B1^-1 K1) pre-smoothing steps, exact
    # dense correction of the level-1 residual, k post-smoothing steps (symmetric => a fixed SPD operator); cost per
    # outer iteration: 1 fine SpMV + (2k [+1 residual]) SpMVs with K1 + ONE dense GEMV (what the additive method has too)
    B1 = lambda c: np.einsum("bij,bj->bi", B1inv, c.reshape(-1, 6)).reshape(-1) * live
    v = rng.standard_normal(6 * S1) * live
    for _ in range(30):
        w = B1(K1 @ v); lb = np.linalg.norm(w) / np.linalg.norm(v); v = w / np.linalg.norm(w)
    lb *= 1.1
    def cheb_smooth(c, y, k, lmax, ratio):
        # k Chebyshev steps on K1 y = c preconditioned by B1, eigenvalue window [lmax/ratio, lmax], from the iterate y
        lmin = lmax / ratio
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        r = c - K1 @ y if y.any() else c.copy()
        sigma = theta / delta; rho = 1.0 / sigma
        d = B1(r) / theta
        for i in range(k):
            y = y + d
            if i + 1 < k:
                r = r - K1 @ d
                rho_new = 1.0 / (2.0 * sigma - rho)
                d = rho_new * rho * d + (2.0 * rho_new / delta) * B1(r)
                rho = rho_new
        return y
    for k, ratio in ((1, 4.0), (2, 6.0), (3, 10.0)):
        def vcycle(c, k=k, ratio=ratio):
            c = c * live
            y = cheb_smooth(c, np.zeros_like(c), k, lb, ratio)
            r = c - K1 @ y
            y = y + P2 @ (E3inv @ (P2.T @ r))
            # post-smoothing = the adjoint of the pre-smoothing: run the same polynomial on the new residual
            r = c - K1 @ y
            y = y + cheb_smooth(r, np.zeros_like(c), k, lb, ratio)
            return y * live
        _, itv = pcg(Kff, b, lambda r: jac(r) + Z1 @ vcycle(Z1.T @ r), rtol, 20000)
        nk1 = 2 * k + 2 * (k - 1)      # K1 products: residual before the dense level, residual before post-smoothing, k-1 inside each smoother
        cost = 1.0 + nk1 * nnzb1 * 36 / (bs.indices.size * 9)
        print(f"   level 1-2 V-cycle, {k} Chebyshev(B1) pre/post steps (lmax {lb:.2f}, ratio {ratio:.0f}): {itv} outer iterations x {cost:.2f} = {itv * cost:.0f}", flush=True)
    # the same with the level-1 term alone (no dense level): how much each level buys
    _, it1 = pcg(Kff, b, lambda r: jac(r) + Z1 @ np.einsum("bij,bj->bi", B1inv, (Z1.T @ r).reshape(-1, 6)).reshape(-1), rtol, 20000)
    print(f"   block-Jacobi + level-1 block-Jacobi only: {it1} iterations", flush=True)


if __name__ == "__main__":
    sizes = tuple(int(x) for x in sys.argv[1].split("x")) if len(sys.argv) > 1 else (20, 4, 4)
    deg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    nps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    S2 = int(sys.argv[4]) if len(sys.argv) > 4 else 32
    run(sizes, deg, nps, S2, boxes=len(sys.argv) > 5 and sys.argv[5] == "boxes")
