"""Multi-GPU driver pieces shared by bench.py and the N>1 tests: one process per GPU
(torchrun), torch.distributed for rendezvous / barriers / max-over-ranks, the library's own NCCL
communicator (created from a unique id broadcast through torch.distributed) for the data path."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


from meshfem_b200.distributed import broadcast_bytes, local_problem  # noqa: E402,F401
from meshfem_b200 import distributed as _dist_mod  # noqa: E402


def make_handle(mfem, dist, world, rank, local_rank, p, D, comm_parent=None, **options):
    """Handle with communicator, local mesh, interface and material (meshfem_b200.distributed)."""
    return _dist_mod.make_handle(dist, world, rank, local_rank, p, D, comm_parent=comm_parent, **options)


def max_over_ranks(dist, value, device):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, value, device):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def run_multi_gpu(args, dist, world, rank, local_rank, name, grid, deg, mat, hbm_peak, peak_src):
    import torch
    import meshfem_b200
    import workloads as wl
    from bench import METRIC, UNIT, RTOL, ClockSampler, workload_name, pinned_copy, precond_name

    device = torch.device("cuda", local_rank)
    m = wl.grid_femmesh(grid, deg)
    D = wl.material(mat)
    fixed, vals, f = wl.cantilever_inputs(m)
    n_elems_total = m.num_elements
    p, lfixed, lvals, lf = local_problem(m, fixed, vals, f, world, rank)
    del m
    nodes_p, _k1 = pinned_copy(p.nodes)
    elems_p, _k2 = pinned_copy(p.elem_nodes)
    f_p, _k3 = pinned_copy(lf)
    p.nodes, p.elem_nodes = nodes_p, elems_p

    sampler = ClockSampler(local_rank)
    opts = {"coarse_aggregates": args.coarse_aggregates} if getattr(args, "coarse_aggregates", 0) else {}
    h = make_handle(meshfem_b200, dist, world, rank, local_rank, p, D, **opts)
    h.assemble()
    h.fix_variables(lfixed, lvals)
    nb, nnzb = h.bsr_sizes()

    def step():
        h.reset_timers()
        h.assemble()
        _, info = h.solve(f_p, rtol=RTOL, return_info=True)
        return h.timer("Assemble System"), info[0]["seconds"], info[0]["iterations"], info[0]["rel_residual"], h.launch_count()

    for _ in range(args.warmup):
        step()
    dist.barrier(); torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    asm_s = solve_s = 0.0
    iters = launches = 0
    relres = None
    for _ in range(args.steps):
        a, s, it, relres, nl = step()
        asm_s += a; solve_s += s; iters += it; launches += nl
    torch.cuda.synchronize(); dist.barrier()
    wall = time.perf_counter() - t0
    spmv_s = h.time_spmv(20)
    clocks = sampler.stop() if rank == 0 else None
    # device time of the step = max over ranks (CUDA-event timers of each rank's stream)
    asm_max = max_over_ranks(dist, asm_s, device)
    solve_max = max_over_ranks(dist, solve_s, device)
    spmv_max = max_over_ranks(dist, spmv_s, device)
    nnzb_tot = sum_over_ranks(dist, nnzb, device)
    nb_tot = sum_over_ranks(dist, nb, device)
    launches_tot = sum_over_ranks(dist, launches, device)

    # end-to-end: fresh handle from host buffers on every rank.  The NCCL communicator is process-level
    # state (like a torch process group): the e2e handles borrow the one created above.
    def e2e_step():
        dist.barrier()
        t = time.perf_counter()
        hh = make_handle(meshfem_b200, dist, world, rank, local_rank, p, D, comm_parent=h, **opts)
        hh.assemble()
        hh.fix_variables(lfixed, lvals)
        u = hh.solve(f_p, rtol=RTOL)
        tip = float(u.reshape(-1, 3)[:, 1].min())
        hh.close()
        torch.cuda.synchronize(); dist.barrier()
        return time.perf_counter() - t, tip
    if args.warmup > 0:
        e2e_step()
    e2e_s, tip = e2e_step()
    e2e_max = max_over_ranks(dist, e2e_s, device)
    tip_min = -max_over_ranks(dist, -tip, device)
    h.close()
    h2d = sum_over_ranks(dist, nodes_p.nbytes + elems_p.nbytes + f_p.nbytes + lfixed.nbytes + lvals.nbytes, device)
    d2h = sum_over_ranks(dist, f_p.nbytes, device)

    if rank == 0:
        dev_s = asm_max + solve_max
        spmv_bytes = nnzb_tot * 76 + nb_tot * 52           # whole-job algorithmic bytes of one distributed SpMV
        out = {
            "metric": METRIC, "value": args.steps * n_elems_total / dev_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(name, grid, deg, mat), "elements": n_elems_total, "rtol": RTOL,
                       "partition": f"{world} x-slabs of elements, shared interface DoFs, NCCL send/recv sum-exchange + all-reduced dots",
                       "preconditioner": precond_name(getattr(args, "coarse_aggregates", 0)), "l2_policy": "inputs larger than L2"},
            "assembly_elements_per_s": args.steps * n_elems_total / asm_max, "pcg_iters_per_s": iters / solve_max,
            "pcg_iterations_per_solve": iters / args.steps, "pcg_rel_residual": relres,
            "assembly_ms": 1e3 * asm_max / args.steps, "solve_ms": 1e3 * solve_max / args.steps,
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "roofline": {"bound": "hbm", "kernel": "k_bsr_spmv (PCG SpMV, per-rank local part, max over ranks)",
                         "achieved": spmv_bytes / spmv_max / 1e9, "peak": hbm_peak * world, "unit": "GB/s",
                         "frac": spmv_bytes / spmv_max / 1e9 / (hbm_peak * world), "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": spmv_bytes, "seconds_per_launch": spmv_max},
            "cpu_baseline": None,
            "e2e": {"value": n_elems_total / e2e_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "seconds_per_step": e2e_max, "steps": 1, "min_uy": tip_min},
            "gpu_launches": int(launches_tot), "clocks": clocks,
        }
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
