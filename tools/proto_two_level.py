"""CPU prototype (numpy/scipy; no GPU): how many PCG iterations would a two-level preconditioner save on the
cantilever workloads?  Level 1 = the block-Jacobi the device PCG uses today; level 2 = a coarse space of per-aggregate
rigid-body modes (6 per aggregate in 3D: the tentative prolongator of smoothed-aggregation AMG), aggregates = contiguous
runs of the Morton-ordered DoFs (what the device numbering already provides).  Variants: additive coarse correction
and deflation (coarse-projected CG).  Prints iteration counts; DESIGN.md section 8 quotes them."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import scipy.sparse as sp
import scipy.linalg as sla
import meshfem_oracle as orc
from util import cantilever_problem


def morton_order(P):
    """Order of the points along a Morton curve (21 bits per axis)."""
    lo, hi = P.min(0), P.max(0)
    q = np.floor((P - lo) / (hi - lo).max() * (2 ** 20 - 1)).astype(np.uint64)
    def spread(x):
        x = x & np.uint64(0x1fffff)
        x = (x | (x << np.uint64(32))) & np.uint64(0x1f00000000ffff)
        x = (x | (x << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
        x = (x | (x << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
        x = (x | (x << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
        x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
        return x
    key = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
    return np.argsort(key, kind="stable")


def pcg(K, b, apply_M, rtol, maxit, project=None):
    x = np.zeros_like(b); r = b.copy()
    if project is not None:
        r = project(r)
    z = apply_M(r); p = z.copy(); rz = r @ z; bb = r @ r
    for it in range(1, maxit + 1):
        Ap = K @ p
        if project is not None:
            Ap = project(Ap)
        a = rz / (p @ Ap)
        x += a * p; r -= a * Ap
        if r @ r <= rtol * rtol * bb:
            return x, it
        z = apply_M(r); rzn = r @ z
        p = z + (rzn / rz) * p; rz = rzn
    return x, maxit


def run(sizes, deg, aggregates, rtol=1e-8, affine=False):
    sim, fixed, vals, f = cantilever_problem(3, deg, sizes)
    K = sim.stiffness().tocsr()
    n = K.shape[0]; N = 3
    free = np.ones(n, bool); free[fixed] = False
    mask = sp.diags(free.astype(float))
    Kff = (mask @ K @ mask + sp.diags((~free).astype(float))).tocsr()
    b = f.reshape(-1) * free
    bs = Kff.tobsr((N, N)); bs.sort_indices()
    nd = n // N
    Minv = np.zeros((nd, N, N))
    for i in range(nd):
        cols = bs.indices[bs.indptr[i]:bs.indptr[i + 1]]
        Minv[i] = np.linalg.inv(bs.data[bs.indptr[i] + np.searchsorted(cols, i)])
    jac = lambda r: np.einsum("bij,bj->bi", Minv, r.reshape(-1, N)).reshape(-1)
    t = time.time()
    _, it0 = pcg(Kff, b, jac, rtol, 20000)
    print(f"grid {sizes} deg {deg}: {nd} nodes, block-Jacobi PCG {it0} iterations ({time.time() - t:.1f}s)", flush=True)
    X = sim.mesh.nodes
    order = morton_order(X)
    for S in aggregates:
        agg = np.empty(nd, dtype=np.int64)
        agg[order] = (np.arange(nd) * S) // nd
        # rigid-body modes per aggregate (centred at the aggregate's centroid for conditioning), masked on fixed variables
        cen = np.zeros((S, 3)); np.add.at(cen, agg, X); cen /= np.bincount(agg, minlength=S)[:, None]
        Y = X - cen[agg]
        nm = 12 if affine else 6
        R = np.zeros((nd, N, nm))
        R[:, 0, 0] = R[:, 1, 1] = R[:, 2, 2] = 1.0
        if affine:      # all affine displacement fields: (1, x, y, z) per component
            scale = np.abs(Y).max()
            for c in range(3):
                for a in range(3):
                    R[:, c, 3 + 3 * c + a] = Y[:, a] / scale
        else:
            R[:, 1, 3], R[:, 2, 3] = -Y[:, 2], Y[:, 1]
            R[:, 0, 4], R[:, 2, 4] = Y[:, 2], -Y[:, 0]
            R[:, 0, 5], R[:, 1, 5] = -Y[:, 1], Y[:, 0]
        rows = np.repeat(np.arange(n), nm)
        cols = (nm * np.repeat(agg, N)[:, None] + np.arange(nm)[None, :]).reshape(-1)
        Z = sp.csr_matrix((R.reshape(-1) * np.repeat(free, nm), (rows, cols)), shape=(n, nm * S))
        E = (Z.T @ Kff @ Z).toarray()
        # aggregates entirely inside the clamped face have zero columns: regularise those
        dead = np.abs(E).sum(axis=1) == 0
        E[dead, dead] = 1.0
        E += 1e-12 * np.trace(E) / E.shape[0] * np.eye(E.shape[0])
        cho = sla.cho_factor(E)
        coarse = lambda r: Z @ sla.cho_solve(cho, Z.T @ r)
        _, it_add = pcg(Kff, b, lambda r: jac(r) + coarse(r), rtol, 20000)
        # deflation: P = I - K Z E^-1 Z^T ; solve P K x~ = P b with PCG, x = Z E^-1 Z^T b + P^T x~
        KZ = Kff @ Z
        proj = lambda v: v - KZ @ sla.cho_solve(cho, Z.T @ v)
        xt, it_def = pcg(Kff, b, jac, rtol, 20000, project=proj)
        x = coarse(b) + (xt - Z @ sla.cho_solve(cho, KZ.T @ xt))
        res = np.linalg.norm(b - Kff @ x) / np.linalg.norm(b)
        print(f"   {S:5d} aggregates ({nd / S:7.1f} nodes each): additive {it_add:5d} it ({it0 / it_add:4.1f}x)   "
              f"deflated {it_def:5d} it ({it0 / it_def:4.1f}x, true rel. residual {res:.1e})", flush=True)


if __name__ == "__main__":
    sizes = tuple(int(x) for x in sys.argv[1].split("x")) if len(sys.argv) > 1 else (20, 4, 4)
    deg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    aggs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [8, 32, 128]
    run(sizes, deg, aggs, affine=len(sys.argv) > 4 and sys.argv[4] == "affine")
