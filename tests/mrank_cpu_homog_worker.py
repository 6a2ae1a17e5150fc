"""World-size-N CPU emulation (gloo) of the distributed periodic homogenization
(meshfem_b200/distributed.py homogenize): the same DoF-based slab partition, shared-DoF lists, owner mask,
pinned variable, exchanged constant-strain loads and volume-form reduction the GPU path uses -- with the
local matrices and loads produced by the CPU oracle and the interface sum-exchange done with
torch.distributed send/recv.  Checks the homogenized tensor against the committed golden one."""
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
    from meshfem_b200 import hostlib
    from meshfem_b200.distributed import _volume_form, local_problem
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    deg, N = 1, 3
    gold = np.load(os.path.join(ROOT, "tests", "golden", "homog_perforated.npz"))
    raw = hostlib.perforated_cell(3, 4, 2)
    info = raw.apply_bc(deg, "", periodic=True)
    m, dfn, nd = info["mesh"], info["dof_for_node"], info["num_dofs"]
    D = orc.isotropic_D(3, 200.0, 0.35)
    p, lfixed, lvals, _ = local_problem(m, info["fixed_vars"], info["fixed_vals"], None, world, rank, dof_for_node=dfn)
    # local mesh in the oracle's terms (only what stiffness / loads / strains need)
    vol, G = orc.embed_simplices(p.nodes[p.elem_nodes[:, :4]])
    loc = type("M", (), {})()
    loc.N, loc.deg, loc.vol, loc.G = 3, deg, vol, G
    loc.elem_nodes, loc.num_elements, loc.num_nodes = p.elem_nodes.astype(np.int64), p.num_elements, p.num_nodes
    ldof = p.dof_for_node
    K = orc.stiffness_matrix(loc, D, ldof, p.num_dofs).tocsr()
    n = N * p.num_dofs
    owned = np.repeat(p.owned.astype(bool), N)

    def exchange_add(v, width):
        v = v.reshape(p.num_dofs, width)
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
    bs = K.tobsr((N, N)); bs.sort_indices()
    diag = np.zeros((p.num_dofs, N, N))
    for i in range(p.num_dofs):
        cols = bs.indices[bs.indptr[i]:bs.indptr[i + 1]]
        diag[i] = bs.data[bs.indptr[i] + np.searchsorted(cols, i)]
    diag = exchange_add(diag.reshape(-1).copy(), N * N).reshape(p.num_dofs, N, N)
    fm = (~free).reshape(p.num_dofs, N)
    for i in range(p.num_dofs):
        for r in range(N):
            if fm[i, r]:
                diag[i, r, :] = 0; diag[i, :, r] = 0; diag[i, r, r] = 1
    Minv = np.linalg.inv(diag)
    apply_M = lambda r: np.einsum("bij,bj->bi", Minv, r.reshape(-1, N)).reshape(-1)
    spmv = lambda x: exchange_add(K @ x, N) * free

    strains, iters = [], []
    for i in range(6):
        # constant-strain load of the local elements, made consistent on the shared DoFs (aux.cu const_strain_load)
        rhs = exchange_add(orc.constant_strain_load(loc, D, -orc.canonical_basis(3, i), ldof, p.num_dofs).reshape(-1).copy(), N)
        b = rhs * free
        x = np.zeros(n); r = b.copy(); z = apply_M(r); pv = z.copy()
        rz = gsum(r[owned] @ z[owned]); bb = gsum(r[owned] @ r[owned])
        its = 0
        while its < 5000 and bb > 0:
            Ap = spmv(pv)
            alpha = rz / gsum(pv[owned] @ Ap[owned])
            x += alpha * pv; r -= alpha * Ap
            z = apply_M(r)
            rzn = gsum(r[owned] @ z[owned]); rr = gsum(r[owned] @ r[owned])
            its += 1
            if rr <= 1e-24 * bb:
                break
            pv = z + (rzn / rz) * pv
            rz = rzn
        iters.append(its)
        w = x.reshape(-1, N)[ldof]                               # dofToNodeField on the local nodes
        strains.append(orc.average_strain_stress(loc, D, w)[0])
    Eh = _volume_form(3, D, vol, strains, float(np.prod(m.bbox_max - m.bbox_min)))
    t = torch.from_numpy(Eh.copy()); dist.all_reduce(t); Eh = t.numpy()
    err = float(np.abs(Eh - gold["Eh_deg1"]).max() / np.abs(Eh).max())
    nowned = gsum(float(p.owned.sum()))
    if rank == 0:
        print(f"MRANK_CPU_HOMOG world={world} iters={iters} err={err:.3e} owned_total={int(nowned)} dofs={nd}", flush=True)
    assert err < 1e-8 and int(nowned) == nd
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
