"""Multi-GPU two-level preconditioner check, run under torchrun on a box with >= 2 GPUs:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/mgpu_two_level_check.py
Every rank solves its slab of a cantilever twice -- block-Jacobi PCG, then with coarse_aggregates on (owner-based
aggregates, all-reduced coarse matrix and coarse residuals; csrc/coarse.inl, multi-GPU variant) -- and compares both
with the CPU oracle's direct solve (rel L2 <= 1e-8); the two-level run must need clearly fewer iterations
(tests/mrank_cpu_worker.py, the gloo emulation of the same algorithm: 581 -> 167 / 143 on 2 / 3 ranks)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import meshfem_b200
    import meshfem_oracle as orc
    import workloads as wl
    from multi_gpu import local_problem, make_handle, max_over_ranks
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl")
    device = torch.device("cuda", local_rank)
    ok = True
    for grid, deg, mat, aggregates in [((20, 4, 4), 2, "iso", 128), ((24, 6, 6), 1, "ortho", 64)]:
        m = wl.grid_femmesh(grid, deg)
        D = wl.material(mat)
        fixed, vals, f = wl.cantilever_inputs(m)
        V, T = orc.grid_simplices(list(grid))
        sim = orc.Simulator(3, deg, V, T); sim.set_material(D)
        u_ref = orc.solve_fixed(sim.stiffness(), f.reshape(-1), fixed, vals).reshape(-1, 3)
        p, lfixed, lvals, lf = local_problem(m, fixed, vals, f, world, rank)
        h = make_handle(meshfem_b200, dist, world, rank, local_rank, p, D)
        h.assemble()
        h.fix_variables(lfixed, lvals)
        u0, info0 = h.solve(lf, rtol=1e-11, return_info=True)
        h.set_option("coarse_aggregates", aggregates)
        u1, info1 = h.solve(lf, rtol=1e-11, return_info=True)
        h.close()
        ref = u_ref[p.nodes_global]
        e0 = max_over_ranks(dist, float(np.linalg.norm(u0.reshape(-1, 3) - ref) / np.linalg.norm(ref)), device)
        e1 = max_over_ranks(dist, float(np.linalg.norm(u1.reshape(-1, 3) - ref) / np.linalg.norm(ref)), device)
        i0, i1 = info0[0]["iterations"], info1[0]["iterations"]
        if rank == 0:
            print(f"grid {grid} deg {deg}: {world} ranks, block-Jacobi {i0} it (err {e0:.2e}), two-level {i1} it (err {e1:.2e})", flush=True)
        ok = ok and e0 < 1e-8 and e1 < 1e-8 and i1 < 0.7 * i0
    dist.barrier()
    dist.destroy_process_group()
    assert ok
    if rank == 0:
        print("MGPU_TWO_LEVEL_OK", flush=True)


if __name__ == "__main__":
    main()
