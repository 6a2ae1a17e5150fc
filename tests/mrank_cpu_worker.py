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
    grid, deg = tuple(int(v) for v in os.environ.get("MRANK_GRID", "10,2,2").split(",")), 2      # MESHFEM_PARTITIONER=rcb: see Partition.hh
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

    # the PCG's matrix-free operator on this rank's elements (numpy restatement of csrc/matfree.inl, tools/emulate_matrix_free.py):
    # equal to the local matrix, and the identity the N-rank iteration relies on -- sum over ranks of the LOCAL products
    # x_loc.(K_loc x_loc) == x.(K x) over owned DoFs after the interface exchange -- holds with it
    import emulate_matrix_free as emf
    tab = emf.build_chunk_tables(p.elem_nodes.astype(np.int64), p.num_nodes, 64)
    Ke = orc.per_element_stiffness(3, deg, vol, G, D)
    xr = np.cos(0.37 * (np.arange(N)[None, :] + 1) * (p.nodes_global[:, None] + 1.0))        # the same values on every sharer
    y_mf, dot_loc = emf.apply_operator(tab, Ke, xr, N)
    assert np.abs(y_mf.reshape(-1) - K @ xr.reshape(-1)).max() <= 1e-12 * np.abs(y_mf).max()
    y_full = exchange_add(y_mf.reshape(-1).copy(), N)
    lhs, rhs = gsum(dot_loc), gsum(xr.reshape(-1)[owned] @ y_full[owned])
    assert abs(lhs - rhs) <= 1e-12 * abs(rhs), (lhs, rhs)
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

    # ---- multilevel variant, exactly the multi-GPU algorithm of csrc/coarse.inl + solver.cu (enqueue_iteration):
    # nested box grids over the OWNED DoFs of each rank; DoFs shared between ranks take no part in level 1 and hang on
    # their OWNER's large box ("pass-through" slots; owner's large id and centred position sent to the sharers by a
    # sum-exchange in which only the owner contributes); E2 = all-reduce of Z2_loc' K_loc Z2_loc; the dense level is
    # ROW-SPLIT (every rank holds the rows of E2^-1 of its own large boxes, computes its slice of y2, one all-gather);
    # per iteration TWO all-reduces: p.Ap = sum over ranks of the LOCAL products p_loc.(K_loc p_loc) over all local
    # rows, and (r.z, r.r, c2) together.
    import scipy.sparse as sp
    import emulate_multilevel as em
    Sr, M, fine = 8, 6, 10
    S2 = Sr * world
    X = p.nodes
    own_n = p.owned.astype(bool)
    shared_n = np.zeros(p.num_nodes, bool)
    for q_ in p.neighbor_ranks:
        shared_n[p.shared[int(q_)]] = True
    lo, hi = X[own_n].min(0), X[own_n].max(0)
    L = hi - lo
    bgrid = np.array(em.choose_boxes(L, Sr))
    rgrid = np.array(em.choose_refinement(L, bgrid, int(own_n.sum()), fine))
    scale1 = np.where(L > 0, bgrid * rgrid / np.where(L > 0, L, 1.0), 0.0)
    qq = np.clip(np.floor((X - lo) * scale1).astype(np.int64), 0, bgrid * rgrid - 1)
    big = np.zeros(p.num_nodes, np.int64); loc = np.zeros(p.num_nodes, np.int64)
    for k in range(3):
        big = big * bgrid[k] + qq[:, k] // rgrid[k]
        loc = loc * rgrid[k] + qq[:, k] % rgrid[k]
    Rr = int(np.prod(rgrid)); nBoxes = int(np.prod(bgrid))
    assert Rr > 1 and nBoxes <= Sr
    S1 = nBoxes * Rr
    aggBase = rank * Sr
    elig = own_n & ~shared_n                                 # level 1: owned, not shared
    T = np.zeros((p.num_nodes, 4))
    cen2 = np.zeros((Sr, 4))
    np.add.at(cen2, big[own_n], np.hstack([X[own_n], np.ones((int(own_n.sum()), 1))]))
    c2pos = cen2[:, :3] / np.maximum(cen2[:, 3:], 1)
    T[own_n, 0] = aggBase + big[own_n] + 1
    T[own_n, 1:] = X[own_n] - c2pos[big[own_n]]
    T = exchange_add(T.reshape(-1).copy(), 4).reshape(p.num_nodes, 4)
    agg2 = np.rint(T[:, 0]).astype(np.int64) - 1
    assert agg2.min() >= 0 and agg2.max() < S2
    Y2 = T[:, 1:]
    slot_small = big * Rr + loc
    cen1 = np.zeros((S1, 4))
    np.add.at(cen1, slot_small[elig], np.hstack([X[elig], np.ones((int(elig.sum()), 1))]))
    has1 = cen1[:, 3] > 0
    shift = np.where((has1 & (cen2[np.arange(S1) // Rr, 3] > 0))[:, None],
                     cen1[:, :3] / np.maximum(cen1[:, 3:], 1) - c2pos[np.arange(S1) // Rr], 0.0)
    slot = np.where(elig, slot_small, S1 + agg2)
    Y1 = np.where(elig[:, None], Y2 - shift[np.minimum(slot_small, S1 - 1)], Y2)
    n1 = S1 + S2
    fm6 = np.repeat(free, M)
    rows = np.repeat(np.arange(n), M)
    P1 = sp.csr_matrix((em.rigid(Y1).reshape(-1) * fm6, (rows, (M * np.repeat(slot, N)[:, None] + np.arange(M)[None, :]).reshape(-1))),
                       shape=(n, M * n1))
    # P2: level-1 slots -> large boxes (global ids); small slots through the shift, pass-through slots by identity
    blocks = np.zeros((n1, M, M))
    for mm in range(M):
        e = np.zeros((S1, M)); e[:, mm] = 1.0
        blocks[:S1, :, mm] = em.shift_prolong(shift, e)
    blocks[S1:] = np.eye(M)
    parent = np.concatenate([aggBase + np.arange(S1) // Rr, np.arange(S2)])
    P2 = sp.csr_matrix((blocks.reshape(-1), (np.repeat(np.arange(M * n1), M), (M * np.repeat(parent, M)[:, None] + np.arange(M)[None, :]).reshape(-1))),
                       shape=(M * n1, M * S2))
    Z2 = (P1 @ P2).tocsr()
    # the rows of Z2 must agree on all sharers: the large boxes' rigid modes at the owner's centred positions
    rows2 = np.repeat(np.arange(n), M)
    Z2direct = sp.csr_matrix((em.rigid(Y2).reshape(-1) * fm6, (rows2, (M * np.repeat(agg2, N)[:, None] + np.arange(M)[None, :]).reshape(-1))),
                             shape=(n, M * S2))
    assert abs(Z2 - Z2direct).max() < 1e-12
    Et = torch.from_numpy((Z2.T @ K @ Z2).toarray())
    dist.all_reduce(Et)
    E = Et.numpy()
    d = np.diag(E).copy()
    E[np.diag_indices_from(E)] = np.where(d == 0.0, 1.0, d * (1 + 1e-8))
    EinvRows = np.linalg.inv(E)[aggBase * M:(aggBase + Sr) * M]           # row-split dense level
    K1 = (P1.T @ K @ P1).tobsr((M, M)); K1.sort_indices()
    B1inv = np.zeros((n1, M, M))
    for s_ in range(S1):
        cols = K1.indices[K1.indptr[s_]:K1.indptr[s_ + 1]]
        k = np.searchsorted(cols, s_)
        if k < cols.size and cols[k] == s_:
            blk = K1.data[K1.indptr[s_] + k]
            B1inv[s_] = em.dropping_cholesky_inverse(0.5 * (blk + blk.T))

    def precond(rv):
        """(z, r.z, r.r) with ONE all-reduce of (r.z_local, r.r_local, c2) and one all-gather of y2."""
        zB = apply_M(rv)
        c1 = (P1.T @ (rv * owned)).reshape(n1, M)
        y1 = np.einsum("sab,sb->sa", B1inv, c1)
        red = np.concatenate([[rv[owned] @ zB[owned] + float((c1 * y1).sum()), rv[owned] @ rv[owned]], P2.T @ c1.reshape(-1)])
        t = torch.from_numpy(red); dist.all_reduce(t)
        red = t.numpy()
        c2v = red[2:]
        y2loc = torch.from_numpy(EinvRows @ c2v)
        parts = [torch.zeros_like(y2loc) for _ in range(world)]
        dist.all_gather(parts, y2loc)
        y2 = np.concatenate([x_.numpy() for x_ in parts])
        qv = y1.reshape(-1) + P2 @ y2
        return zB + P1 @ qv, red[0] + float(c2v @ y2), red[1]

    x = np.zeros(n); r = b.copy()
    z, rz, _ = precond(r)
    pvec = z.copy()
    its2 = 0
    while its2 < 5000:
        Aloc = (K @ pvec) * free                               # local product, before the exchange
        pAp = gsum(float(pvec @ Aloc))                         # all local rows, no owner mask: K = sum of the ranks' K_loc
        Ap = exchange_add(Aloc.copy(), N)
        alpha = rz / pAp
        x += alpha * pvec; r -= alpha * Ap
        z, rzn, rr = precond(r)
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
    assert t2.item() < 1e-8 and its2 < 0.6 * its, (t2.item(), its, its2)
    t = torch.tensor([err], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nowned = gsum(float(p.owned.sum()))
    if rank == 0:
        print(f"MRANK_CPU world={world} iters={its} err={t.item():.3e} two_level_iters={its2} two_level_err={t2.item():.3e} owned_total={int(nowned)} nodes={m.num_nodes}", flush=True)
    assert t.item() < 1e-8 and int(nowned) == m.num_nodes
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
