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
    V, T = orc.grid_simplices(list(grid))
    sim = orc.Simulator(3, deg, V, T); sim.set_material(D)
    u_ref = orc.solve_fixed(sim.stiffness(), f.reshape(-1), fixed, vals).reshape(-1, N)[p.nodes_global]
    err = float(np.linalg.norm(u - u_ref) / np.linalg.norm(u_ref))
    t = torch.tensor([err], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nowned = gsum(float(p.owned.sum()))
    if rank == 0:
        print(f"MRANK_CPU world={world} iters={its} err={t.item():.3e} owned_total={int(nowned)} nodes={m.num_nodes}", flush=True)
    assert t.item() < 1e-8 and int(nowned) == m.num_nodes
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
