"""World-size-N CPU emulation (gloo) of the multi-GPU solver's host-side logic: the same slab
partition, shared-DoF lists, owner mask and exchange order the GPU path uses
(meshfem_b200/csrc/comm.cu), with the local matrices assembled by the CPU oracle and the interface
sum-exchange done with torch.distributed send/recv.  Checks the distributed block-Jacobi PCG
against the oracle's direct solve of the whole problem."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import meshfem_oracle as orc
    import workloads as wl
    from multi_gpu import local_problem
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    grid, deg = (10, 2, 2), 2
    m = wl.grid_femmesh(grid, deg)
    D = wl.material("ortho")
    fixed, vals, f = wl.cantilever_inputs(m)
    vals = vals + 1e-3 * np.cos(np.arange(vals.size))
    p, lfixed, lvals, lf = local_problem(m, fixed, vals, f, world, rank)
    N = 3
    # local K over this rank's elements only: partial sums on interface rows
    lm = orc.build_mesh(3, deg, p.nodes[:int(p.elem_nodes[:, :4].max()) + 1], p.elem_nodes[:, :4]) if False else None
    vol, G = orc.embed_simplices(p.nodes[p.elem_nodes[:, :4]])
    loc = type("M", (), {})()
    loc.N, loc.deg, loc.vol, loc.G, loc.elem_nodes, loc.num_elements, loc.num_nodes = 3, deg, vol, G, p.elem_nodes.astype(np.int64), p.num_elements, p.num_nodes
    K = orc.stiffness_matrix(loc, D).tocsr()
    n = N * p.num_nodes
    owned = np.repeat(p.owned.astype(bool), N)

    def exchange_add(v, width):
        v = v.reshape(p.num_nodes, width)
        reqs, recvs = [], []
        for q in p.neighbor_ranks:
            idx = p.shared[int(q)]
            send = torch.from_numpy(np.ascontiguousarray(v[idx]))
            recv = torch.zeros_like(send)
            reqs.append(dist.isend(send, int(q))); reqs.append(dist.irecv(recv, int(q)))
            recvs.append((idx, recv))
        for r in reqs:
            r.wait()
        for idx, recv in recvs:
            v[idx] += recv.numpy()
        return v.reshape(-1)

    def gsum(x):
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    free = np.ones(n, bool); free[lfixed] = False
    ufix = np.zeros(n); ufix[lfixed] = lvals
    # block-Jacobi on the COMPLETED diagonal blocks
    bs = K.tobsr((N, N)); bs.sort_indices()
    diag = np.zeros((p.num_nodes, N, N))
    for i in range(p.num_nodes):
        cols = bs.indices[bs.indptr[i]:bs.indptr[i + 1]]
        diag[i] = bs.data[bs.indptr[i] + np.searchsorted(cols, i)]
    diag = exchange_add(diag.reshape(-1).copy(), N * N).reshape(p.num_nodes, N, N)
    fm = (~free).reshape(p.num_nodes, N)
    for i in range(p.num_nodes):
        for r in range(N):
            if fm[i, r]:
                diag[i, r, :] = 0; diag[i, :, r] = 0; diag[i, r, r] = 1
    Minv = np.linalg.inv(diag)
    apply_M = lambda r: np.einsum("bij,bj->bi", Minv, r.reshape(-1, N)).reshape(-1)
    spmv = lambda x: exchange_add(K @ x, N) * free
    b = (lf.reshape(-1) - exchange_add(K @ ufix, N)) * free
    x = np.zeros(n); r = b.copy(); z = apply_M(r); pvec = z.copy()
    rz = gsum(r[owned] @ z[owned]); bb = gsum(r[owned] @ r[owned])
    its = 0
    while its < 5000:
        Ap = spmv(pvec)
        alpha = rz / gsum(pvec[owned] @ Ap[owned])
        x += alpha * pvec; r -= alpha * Ap
        z = apply_M(r)
        rzn = gsum(r[owned] @ z[owned]); rr = gsum(r[owned] @ r[owned])
        its += 1
        if rr <= 1e-24 * bb:
            break
        pvec = z + (rzn / rz) * pvec
        rz = rzn
    u = (x + ufix).reshape(-1, N)

    # ---- two-level variant, exactly the multi-GPU algorithm of csrc/coarse.inl (build_coarse_indexed_impl): box aggregates
    # over the OWNED DoFs of each rank, owner's
    # aggregate id and centred position sent to the sharers by a sum-exchange in which only the owner contributes,
    # E = all-reduce of Z_loc' K_loc Z_loc, restriction over owned DoFs + all-reduce, replicated coarse solve,
    # prolongation on every local DoF, r.z corrected by c.y after its all-reduce.
    Sr, M = 8, 6
    S = Sr * world
    X = p.nodes
    own_n = p.owned.astype(bool)
    # near-cubic boxes over the bounding box of the OWNED nodes, at most Sr of them (coarse_choose_boxes / k_coarse_box_agg)
    lo, hi = X[own_n].min(0), X[own_n].max(0)
    L = hi - lo
    h = (np.prod(L[L > 0]) / Sr) ** (1.0 / (L > 0).sum())
    bx = [int(max(1, np.floor(l / h + 0.5))) if l > 0 else 1 for l in L]
    while np.prod(bx) > Sr:
        k = int(np.argmax(bx))
        if bx[k] == 1:
            break
        bx[k] -= 1
    scale = np.where(L > 0, np.array(bx) / np.where(L > 0, L, 1.0), 0.0)
    q = np.clip(np.floor((X - lo) * scale).astype(np.int64), 0, np.array(bx) - 1)
    box = (q[:, 0] * bx[1] + q[:, 1]) * bx[2] + q[:, 2]
    agg = np.where(own_n, rank * Sr + box, -1)
    T = np.zeros((p.num_nodes, 4))
    cen = np.zeros((Sr, 4))
    np.add.at(cen, agg[own_n] - rank * Sr, np.hstack([X[own_n], np.ones((int(own_n.sum()), 1))]))
    T[own_n, 0] = agg[own_n] + 1
    T[own_n, 1:] = X[own_n] - (cen[:, :3] / np.maximum(cen[:, 3:], 1))[agg[own_n] - rank * Sr]
    T = exchange_add(T.reshape(-1).copy(), 4).reshape(p.num_nodes, 4)
    agg = np.rint(T[:, 0]).astype(np.int64) - 1
    assert agg.min() >= 0 and agg.max() < S
    Y = T[:, 1:]
    R = np.zeros((p.num_nodes, N, M))
    R[:, 0, 0] = R[:, 1, 1] = R[:, 2, 2] = 1.0
    R[:, 1, 3], R[:, 2, 3] = -Y[:, 2], Y[:, 1]
    R[:, 0, 4], R[:, 2, 4] = Y[:, 2], -Y[:, 0]
    R[:, 0, 5], R[:, 1, 5] = -Y[:, 1], Y[:, 0]
    import scipy.sparse as sp
    rows = np.repeat(np.arange(n), M)
    cols = (M * np.repeat(agg, N)[:, None] + np.arange(M)[None, :]).reshape(-1)
    Z = sp.csr_matrix((R.reshape(-1) * np.repeat(free, M), (rows, cols)), shape=(n, M * S))
    Et = torch.from_numpy((Z.T @ K @ Z).toarray())
    dist.all_reduce(Et)
    E = Et.numpy()
    d = np.diag(E).copy()
    E[np.diag_indices_from(E)] = np.where(d == 0.0, 1.0, d * (1 + 1e-8))
    Einv = np.linalg.inv(E)
    Zown = sp.diags(owned.astype(float)) @ Z

    def coarse(rv):
        ct = torch.from_numpy(Zown.T @ rv); dist.all_reduce(ct)
        cv = ct.numpy(); yv = Einv @ cv
        return Z @ yv, float(cv @ yv)

    x = np.zeros(n); r = b.copy(); z = apply_M(r)
    rz = gsum(r[owned] @ z[owned]); zc, cy = coarse(r); z = z + zc; rz += cy
    pvec = z.copy()
    its2 = 0
    while its2 < 5000:
        Ap = spmv(pvec)
        alpha = rz / gsum(pvec[owned] @ Ap[owned])
        x += alpha * pvec; r -= alpha * Ap
        z = apply_M(r)
        rzn = gsum(r[owned] @ z[owned]); rr = gsum(r[owned] @ r[owned])
        zc, cy = coarse(r); z = z + zc; rzn += cy
        its2 += 1
        if rr <= 1e-24 * bb:
            break
        pvec = z + (rzn / rz) * pvec
        rz = rzn
    u2 = (x + ufix).reshape(-1, N)
    V, T = orc.grid_simplices(list(grid))
    sim = orc.Simulator(3, deg, V, T); sim.set_material(D)
    u_ref = orc.solve_fixed(sim.stiffness(), f.reshape(-1), fixed, vals).reshape(-1, N)[p.nodes_global]
    err = float(np.linalg.norm(u - u_ref) / np.linalg.norm(u_ref))
    err2 = float(np.linalg.norm(u2 - u_ref) / np.linalg.norm(u_ref))
    t2 = torch.tensor([err2], dtype=torch.float64); dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    assert t2.item() < 1e-8 and its2 < 0.8 * its, (t2.item(), its, its2)
    t = torch.tensor([err], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nowned = gsum(float(p.owned.sum()))
    if rank == 0:
        print(f"MRANK_CPU world={world} iters={its} err={t.item():.3e} two_level_iters={its2} two_level_err={t2.item():.3e} owned_total={int(nowned)} nodes={m.num_nodes}", flush=True)
    assert t.item() < 1e-8 and int(nowned) == m.num_nodes
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
