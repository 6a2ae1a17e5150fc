"""BASELINE config 4: periodic homogenization of the perforated unit cell (n^3 voxels minus the centred
(n/2)^3 block, 24 tets per voxel, quadratic tets, isotropic base E=200 nu=0.35), six cell problems solved as one
batched PCG, on 1..N GPUs:
  python tools/homog_bench.py [n] [degree]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/homog_bench.py 64 2
Prints one JSON line (rank 0): sizes, iterations, device seconds of the batched solve (max over ranks), wall
seconds of the whole homogenization, the homogenized tensor and its symmetry / cubic-symmetry residuals."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshfem_b200 import build as mb, distributed, hostlib  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    deg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    batch = int(os.environ.get("BATCH_RHS", "1"))
    coarse = int(os.environ.get("COARSE", "-1"))          # -1: automatic multilevel preconditioner (systems one after the other), 0: batched block-Jacobi PCG
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl")
        dist = dist_mod
    mb.build_all()
    E, nu = 200.0, 0.35
    lam, mu = nu * E / ((1 + nu) * (1 - 2 * nu)), E / (2 + 2 * nu)
    D = np.zeros((6, 6)); D[:3, :3] = lam; D[np.arange(3), np.arange(3)] = lam + 2 * mu; D[np.arange(3, 6), np.arange(3, 6)] = mu
    t0 = time.perf_counter()
    raw = hostlib.perforated_cell(3, n, n // 2)
    t_mesh = time.perf_counter() - t0
    t0 = time.perf_counter()
    Eh, extra = distributed.homogenize(raw, deg, D, dist=dist, local_rank=local_rank, rtol=1e-8, return_fields=True, batch_rhs=batch, coarse_aggregates=coarse)
    wall = time.perf_counter() - t0
    solve_s = sum(s["seconds"] for s in extra["solves"])
    if dist is not None:
        import torch
        t = torch.tensor([solve_s, wall], dtype=torch.float64, device=torch.device("cuda", local_rank))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        solve_s, wall = float(t[0]), float(t[1])
    if rank == 0:
        m = extra["mesh"]
        out = {
            "workload": f"cfg4: {n}^3 voxel cell minus centred {n // 2}^3 block, degree {deg}, periodic, {6} cell problems",
            "n_gpus": world, "elements": int(m.num_elements), "nodes": int(m.num_nodes), "dofs": 3 * int(extra["dof_for_node"].max() + 1),
            "batched": bool(batch) and coarse == 0, "coarse_aggregates": coarse, "iterations": [s["iterations"] for s in extra["solves"]],
            "solve_device_s": solve_s, "homogenize_wall_s": wall, "mesh_generation_s": t_mesh,
            "elements_per_s_solve": 6 * m.num_elements / solve_s,
            "Eh_diag": [float(Eh[i, i]) for i in range(6)], "Eh_01": float(Eh[0, 1]),
            "asymmetry": float(np.abs(Eh - Eh.T).max() / np.abs(Eh).max()),
            "cubic_symmetry_residual": float(max(abs(Eh[0, 0] - Eh[1, 1]), abs(Eh[0, 0] - Eh[2, 2]), abs(Eh[3, 3] - Eh[4, 4]),
                                                 abs(Eh[3, 3] - Eh[5, 5]), abs(Eh[0, 1] - Eh[0, 2]), abs(Eh[0, 1] - Eh[1, 2])) / abs(Eh[0, 0])),
        }
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
