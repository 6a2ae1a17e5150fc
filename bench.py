#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 linear-elasticity assemble-and-solve path.

Metric (BASELINE.json): assembly + solve throughput in elements/s on the synthetic `grid -t`
cantilever of the named size (default: config 5, 220x44x44 hexes -> 10,222,080 quadratic
tets, 43.96M DoF, isotropic E=200 nu=0.35, examples/cantilever/cantilever.bc), with the PCG
iteration rate and the SpMV roofline beside it.

One "step" = one numeric assembly of K into the cached block-CSR pattern + the set-up of the
preconditioner for the new values (block-Jacobi blocks, coarse matrices) + one PCG solve to rtol
on ||r||/||b||.  `value` times steps with all inputs resident in HBM (CUDA events on the
library's stream, max over ranks); `e2e` times the same work through the C ABI from HOST buffers
on a fresh handle (mesh upload, symbolic phase, assembly, constraints, solve, result download all
inside the timed region).  N GPUs (torchrun, one process per GPU): x-slabs of elements, the
library's own NCCL communicator for the interface exchange and the all-reduces.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg5|cfg3|cfg2|GRID:DEG:MAT]
  python bench.py --impl reference ...     # the reference algorithm on the host cores (oracle port)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402

METRIC = "assembly+solve throughput (elements/s), quadratic-tet elasticity cantilever"
UNIT = "elements/s"
RTOL = 1e-8
# bounded CPU samples of the workload (same cross-section ratio 5:1:1, same material / BCs), with the seconds one
# assemble+solve took on the 16 host threads of the round-1 GPU box: the reference arm picks the largest one
# that keeps `steps + warmup` samples within its time budget
CPU_SAMPLE_GRIDS = [((72, 14, 14), 40.0), ((60, 12, 12), 21.0), ((50, 10, 10), 10.0), ((40, 8, 8), 4.2), ((30, 6, 6), 1.4)]
CPU_ARM_BUDGET_S = 400.0
DIRECT_GRID = (24, 5, 5)          # direct-solver leg: 14,400 quadratic tets, 67,767 DoF (SuperLU needs ~8 s on one core)


def parse_config(s):
    import workloads as wl
    if s in wl.CONFIGS:
        grid, deg, mat = wl.CONFIGS[s]
        return s, tuple(grid), deg, mat
    g, d, m = s.split(":")
    return s, tuple(int(x) for x in g.split("x")), int(d), m


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(cfg, kernel, nnzb):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of `kernel` from the committed
    `ncu --set full` capture of this workload (profiles/traffic.json), or None when there is no capture
    for this exact matrix."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)[cfg][kernel]
        return float(t["dram_bytes_per_launch"]) if int(t["nnz_blocks"]) == int(nnzb) else None
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if r[4 + k].lower().startswith("active"):
                    reasons.add(nme)
        # "under load": samples whose power draw is above the midpoint of the observed range
        if sm:
            thr = 0.5 * (min(power) + max(power))
            loaded = [s for s, p in zip(sm, power) if p >= thr] or sm
            return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                    "samples": len(sm), "power_w_max": max(power)}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(reasons)}


def pinned_copy(a):
    """Copy of `a` in page-locked host memory (torch is plumbing for the allocation only)."""
    try:
        import torch
        t = torch.empty(a.shape, dtype=getattr(torch, str(a.dtype)), pin_memory=True)
        out = t.numpy()
        out[...] = a
        out_ref = (out, t)            # keep the tensor alive
        return out_ref
    except Exception:
        return (np.ascontiguousarray(a), None)


# ---------------------------------------------------------------------------------------------
def cpu_port_sample(grid, deg, mat, rtol, threads=None):
    """The reference algorithm on the host cores, on a bounded sample of the workload:
    perElementStiffness loop nest + serial triplet scatter + sumRepeated (oracle/ref_cpu.cc,
    following LinearElasticity.hh:165-232, 1408-1466 and SparseMatrices.hh:280-374), then a
    block-Jacobi PCG on the same cores standing in for CHOLMOD (not buildable here)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_cpu
    import workloads as wl
    m = wl.grid_femmesh(grid, deg)
    D = wl.material(mat)
    fixed, vals, f = wl.cantilever_inputs(m)
    threads = threads or os.cpu_count() or 1
    A, t_asm = ref_cpu.assemble_upper_csc(3, deg, m.nodes, m.elem_nodes, D, threads=threads)
    u, info = ref_cpu.solve_fixed_pcg(3, A, f, fixed, vals, rtol=rtol, threads=threads)
    t_assemble = t_asm["ke"] + t_asm["scatter"] + t_asm["compress"]
    total = t_assemble + info["seconds"]
    return {
        "value": m.num_elements / total, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": (f"grid {'x'.join(map(str, grid))} -t, degree {deg}, {m.num_elements} elements, same material/BCs; "
                   f"assembly {t_assemble:.2f}s (Ke {t_asm['ke']:.2f} + serial scatter {t_asm['scatter']:.2f} + compress "
                   f"{t_asm['compress']:.2f}), block-Jacobi PCG {info['iters']} it in {info['seconds']:.2f}s "
                   f"(CHOLMOD stand-in; smaller mesh => fewer iterations than the full workload, i.e. favourable to the CPU)"),
        "seconds": total, "elements": m.num_elements, "pcg_iterations": info["iters"],
        "assembly_seconds": t_assemble, "solve_seconds": info["seconds"],
    }


def cpu_direct_sample(deg, mat):
    """Direct-solver leg of the CPU baseline (the reference's solver class: SPSDSystem::solve -> CholmodFactorizer,
    SparseMatrices.hh:2002-2024, 2106-2124) on a size where a direct factorisation finishes in seconds: reference-style
    assembly (oracle/ref_cpu.cc) + sparse factorisation and solve of K_ff.  CHOLMOD is not available; the stand-ins are
    cuSOLVER's HOST sparse Cholesky with METIS nested dissection (cusolverSpDcsrlsvcholHost, reorder = 3: a CPU routine,
    the closest thing to CHOLMOD's NESDIS + Cholesky in this image) when libcusolver loads, and SuperLU (scipy splu,
    MMD on A^T + A, symmetric mode) otherwise / besides.  One thread each."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ctypes
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import ref_cpu
    import workloads as wl
    m = wl.grid_femmesh(DIRECT_GRID, deg)
    D = wl.material(mat)
    fixed, vals, f = wl.cantilever_inputs(m)
    threads = os.cpu_count() or 1
    A, t_asm = ref_cpu.assemble_upper_csc(3, deg, m.nodes, m.elem_nodes, D, threads=threads)
    t_assemble = t_asm["ke"] + t_asm["scatter"] + t_asm["compress"]
    K = (A + sp.triu(A, 1).T).tocsr()
    n = K.shape[0]
    free = np.ones(n, dtype=bool); free[np.asarray(fixed)] = False
    Kff = K[free][:, free].tocsr(); Kff.sort_indices()
    b = np.asarray(f, float).reshape(-1)[free].copy()
    out = {"grid": "x".join(map(str, DIRECT_GRID)), "elements": int(m.num_elements), "dofs": int(Kff.shape[0]),
           "assembly_seconds": round(t_assemble, 3), "legs": {}}
    try:        # cuSOLVER host sparse Cholesky (CPU code path of libcusolver)
        lib = ctypes.CDLL("libcusolver.so.11")
        sparse = ctypes.CDLL("libcusparse.so.12")
        h, descr = ctypes.c_void_p(), ctypes.c_void_p()
        if lib.cusolverSpCreate(ctypes.byref(h)) != 0 or sparse.cusparseCreateMatDescr(ctypes.byref(descr)) != 0:
            raise RuntimeError("cusolverSpCreate failed")
        x = np.zeros_like(b); sing = ctypes.c_int(0)
        ip, ix = Kff.indptr.astype(np.int32), Kff.indices.astype(np.int32)
        t0 = time.perf_counter()
        st = lib.cusolverSpDcsrlsvcholHost(h, ctypes.c_int(Kff.shape[0]), ctypes.c_int(Kff.nnz), descr,
                                           Kff.data.ctypes.data_as(ctypes.c_void_p), ip.ctypes.data_as(ctypes.c_void_p),
                                           ix.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                                           ctypes.c_double(0.0), ctypes.c_int(3), x.ctypes.data_as(ctypes.c_void_p), ctypes.byref(sing))
        dt = time.perf_counter() - t0
        res = float(np.linalg.norm(Kff @ x - b) / np.linalg.norm(b))
        if st != 0 or sing.value != -1 or not res < 1e-6:
            raise RuntimeError(f"status {st}, singularity {sing.value}, residual {res:.1e}")
        out["legs"]["cusolverSp host Cholesky (METIS nested dissection, 1 thread)"] = {
            "factor_solve_seconds": round(dt, 3), "rel_residual": res, "elements_per_s": m.num_elements / (t_assemble + dt)}
    except Exception as e:  # noqa: BLE001
        out["legs"]["cusolverSp host Cholesky"] = {"unavailable": f"{type(e).__name__}: {e}"[:160]}
    t0 = time.perf_counter()
    lu = spla.splu(Kff.tocsc(), permc_spec="MMD_AT_PLUS_A", options=dict(SymmetricMode=True))
    x = lu.solve(b)
    dt = time.perf_counter() - t0
    out["legs"]["SuperLU (scipy splu, MMD_AT_PLUS_A, symmetric mode, 1 thread)"] = {
        "factor_solve_seconds": round(dt, 3), "rel_residual": float(np.linalg.norm(Kff @ x - b) / np.linalg.norm(b)),
        "factor_nnz": int(lu.L.nnz + lu.U.nnz), "elements_per_s": m.num_elements / (t_assemble + dt)}
    out["note"] = ("direct sparse factorisation as the reference does it, at a size where it takes seconds: cost grows like "
                   "n^2 (3D fill), so at the PCG sample's size it would be ~100x slower than the PCG stand-in used for `value`")
    return out


def cpu_sample_grid(n_samples):
    per_step = min(60.0, CPU_ARM_BUDGET_S / max(n_samples, 1))
    for grid, secs in CPU_SAMPLE_GRIDS:
        if secs <= per_step:
            return grid
    return CPU_SAMPLE_GRIDS[-1][0]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name, grid, deg, mat = parse_config(args.config)
    sgrid = cpu_sample_grid(args.steps + args.warmup) if deg == 2 else (50, 10, 10)
    samples = []
    for _ in range(args.warmup):
        cpu_port_sample(sgrid, deg, mat, RTOL)
    for _ in range(args.steps):
        samples.append(cpu_port_sample(sgrid, deg, mat, RTOL))
    secs = sum(s["seconds"] for s in samples)
    elems = sum(s["elements"] for s in samples)
    value = elems / secs
    base = dict(samples[-1]); base["value"] = value
    for k in ("seconds", "elements", "pcg_iterations", "assembly_seconds", "solve_seconds"):
        base.pop(k, None)
    base["pcg"] = {"elements_per_s": value, "grid": "x".join(map(str, sgrid))}
    if not args.no_direct:
        base["direct"] = cpu_direct_sample(deg, mat)
    sample_name = (f"{name} SAMPLE: grid {'x'.join(map(str, sgrid))} -t ({24 * sgrid[0] * sgrid[1] * sgrid[2]} "
                   f"{'quadratic' if deg == 2 else 'linear'} tets), {mat} material, cantilever.bc")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": sample_name, "full_workload": workload_name(name, grid, deg, mat), "rtol": RTOL,
                   "same_config": False,
                   "note": ("reference algorithm (oracle port: perElementStiffness loop nest, serial triplet scatter, sumRepeated; "
                            "block-Jacobi PCG standing in for CHOLMOD) on the host cores.  The full workload needs ~295 GB of "
                            "triplets and hours on this host, so every step runs a bounded SAMPLE of it (the grid named in "
                            "`workload`); throughput in elements/s on the smaller mesh flatters the CPU (fewer PCG iterations).")},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def precond_name(aggregates, fine):
    if not aggregates:
        return "block-Jacobi 3x3"
    s = "multilevel aggregation: block-Jacobi 3x3 + "
    if fine:
        s += f"level-1 rigid modes of ~{fine}-node boxes (6x6 block solves) + "
    s += ("dense level of rigid modes of " + (f"<= {aggregates}" if aggregates > 0 else "automatically many") + " large boxes (additive)")
    return s


def workload_name(name, grid, deg, mat):
    return (f"{name}: grid {'x'.join(map(str, grid))} -t ({24 * grid[0] * grid[1] * grid[2]} "
            f"{'quadratic' if deg == 2 else 'linear'} tets), {mat} material, cantilever.bc")


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = device = None
    # stdout carries ONE JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION prints to stdout) out of it
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
        device = torch.device("cuda", local_rank)
    # one rank builds (a stale stamp would otherwise start N concurrent nvcc / g++ runs into the same files)
    from meshfem_b200 import build as mb
    if rank == 0:
        mb.build_all()
    if dist is not None:
        dist.barrier()
    import meshfem_b200
    import workloads as wl
    from multi_gpu import local_problem, make_handle, max_over_ranks, sum_over_ranks

    def rmax(v):
        return max_over_ranks(dist, v, device) if dist is not None else float(v)

    def rsum(v):
        return sum_over_ranks(dist, v, device) if dist is not None else float(v)

    name, grid, deg, mat = parse_config(args.config)
    hbm_peak, peak_src = peaks()
    t_gen = time.perf_counter()
    m = wl.grid_femmesh(grid, deg)
    D = wl.material(mat)
    fixed, vals, f = wl.cantilever_inputs(m)
    n_elems, n_nodes, npe = m.num_elements, m.num_nodes, m.elem_nodes.shape[1]
    if world > 1:
        p, fixed, vals, f = local_problem(m, fixed, vals, f, world, rank)
        nodes, elem_nodes = p.nodes, p.elem_nodes
    else:
        p, nodes, elem_nodes = None, m.nodes, m.elem_nodes
    del m
    t_gen = time.perf_counter() - t_gen
    nodes_p, _k1 = pinned_copy(nodes)
    elems_p, _k2 = pinned_copy(elem_nodes)
    f_p, _k3 = pinned_copy(np.ascontiguousarray(f))
    u_p, _k4 = pinned_copy(np.zeros_like(np.ascontiguousarray(f)))      # page-locked result buffer, reused by every solve
    if p is not None:
        p.nodes, p.elem_nodes = nodes_p, elems_p
    opts = {"coarse_aggregates": args.coarse_aggregates, "coarse_fine_nodes": args.coarse_fine_nodes, "matrix_free": args.matrix_free}
    if world > 1:
        opts["comm_p2p"] = args.comm_p2p

    def new_handle(parent=None, **over):
        o = dict(opts); o.update(over)
        if world > 1:
            return make_handle(meshfem_b200, dist, world, rank, local_rank, p, D, comm_parent=parent, **o)
        hh = meshfem_b200.Handle(local_rank, **o)
        hh.set_mesh(3, deg, nodes_p, elems_p)
        hh.set_material(D)
        return hh

    sampler = ClockSampler(local_rank)
    # ------------------------------------------------------------------ device-resident steps
    h = new_handle()
    uses_window = world > 1 and h.comm_uses_peer_window()
    h.assemble()                      # symbolic phase (pattern + incidence lists) is cached from here on
    h.fix_variables(fixed, vals)
    nb, nnzb = h.bsr_sizes()
    pattern_s = h.timer("Pattern")

    def step(hh=None, rtol=RTOL):
        hh = hh or h
        hh.reset_timers()
        hh.assemble()
        u, info = hh.solve(f_p, rtol=rtol, max_iters=20000, return_info=True)
        # block-Jacobi blocks + coarse matrices are rebuilt for the new values inside solve(): counted in the step
        setup = max(0.0, hh.timer("Fix Variables")) + max(0.0, hh.timer("Coarse Space"))
        return u, hh.timer("Assemble System"), setup, info[0]["seconds"], info[0]["iterations"], info[0]["rel_residual"], hh.launch_count()

    for _ in range(args.warmup):
        step()
    if dist is not None:
        import torch
        dist.barrier(); torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    wall0 = time.perf_counter()
    asm_s = setup_s = solve_s = 0.0
    iters = 0
    relres, u = None, None
    launches = 0
    for _ in range(args.steps):
        u, a, su, s, it, relres, nl = step()          # reset_timers() restarts the launch counter: nl is per step
        asm_s += a; setup_s += su; solve_s += s; iters += it; launches += nl
    if dist is not None:
        torch.cuda.synchronize(); dist.barrier()
    wall = time.perf_counter() - wall0
    # dominant kernel: the PCG SpMV (the masked + fused-dot variant the iteration launches), timed live on the
    # library's stream (inputs: the 31 GB matrix >> L2)
    spmv_s = h.time_spmv(20)
    # ... and the product the iteration actually launches: the mesh-based operator (csrc/matfree.inl) where option
    # matrix_free selects it (3D quadratic elements by default), its two kernels also timed alone
    op_s, op_parts, mf_used = h.time_operator(20)
    clocks = sampler.stop() if rank == 0 else None
    asm_max, setup_max, solve_max, spmv_max = rmax(asm_s), rmax(setup_s), rmax(solve_s), rmax(spmv_s)
    op_max, op_elem_max, op_gather_max = rmax(op_s), rmax(op_parts[0]), rmax(op_parts[1])
    n_partials = rsum(max(0.0, h.timer("Matrix-free Partials")))       # chunked operator: (chunk, DoF) partial sums; 0 = one slot per (element, node)
    mf_plan_s = rmax(max(0.0, h.timer("Matrix-free Plan")))
    nnzb_tot, nb_tot, launches_tot = rsum(nnzb), rsum(nb), rsum(launches)
    dev_s = rmax(asm_s + setup_s + solve_s)
    coarse_sizes = None
    if args.coarse_aggregates:
        try:
            cz = h.coarse_array("sizes")
            coarse_sizes = {"small_boxes_this_rank": int(cz[0]), "large_boxes": int(cz[1]), "small_per_large": int(cz[2])}
        except Exception:
            coarse_sizes = None

    # ------------------------------------------------------------------ parity evidence at full size
    parity = None
    if not args.no_parity:
        u = np.asarray(u).reshape(-1)
        parity = {"rtol": RTOL}
        # (1) a tighter solve with the same preconditioner: how far is the rtol answer from the converged one
        u12, *_ = step(rtol=1e-12)
        u12 = np.asarray(u12).reshape(-1)
        num, den = rsum(float(np.sum((u - u12) ** 2))), rsum(float(np.sum(u12 ** 2)))
        parity["rel_l2_rtol_1e-8_vs_1e-12"] = float(np.sqrt(num / den))
        # (2) an independent preconditioner: block-Jacobi PCG on the same matrix, full vector
        if args.coarse_aggregates and not args.no_bj_parity:
            h.set_option("coarse_aggregates", 0)
            ubj, info = h.solve(f_p, rtol=RTOL, return_info=True)
            h.set_option("coarse_aggregates", args.coarse_aggregates)
            ubj = np.asarray(ubj).reshape(-1)
            num, den = rsum(float(np.sum((u - ubj) ** 2))), rsum(float(np.sum(ubj ** 2)))
            parity["rel_l2_vs_block_jacobi_pcg"] = float(np.sqrt(num / den))
            parity["block_jacobi_iterations"] = int(info[0]["iterations"])
            parity["block_jacobi_solve_ms"] = 1e3 * rmax(info[0]["seconds"])
        # (2b) an independent operator: the same PCG multiplying with the STORED matrix (block-CSR SpMV) instead of the
        #      mesh-based operator, full vector
        if mf_used:
            h.set_option("matrix_free", 0)
            ust, info = h.solve(f_p, rtol=RTOL, return_info=True)
            h.set_option("matrix_free", args.matrix_free)
            ust = np.asarray(ust).reshape(-1)
            num, den = rsum(float(np.sum((u - ust) ** 2))), rsum(float(np.sum(ust ** 2)))
            parity["rel_l2_vs_stored_matrix_pcg"] = float(np.sqrt(num / den))
            parity["stored_matrix_iterations"] = int(info[0]["iterations"])
            parity["stored_matrix_solve_ms"] = 1e3 * rmax(info[0]["seconds"])
        # (3) true residual of the returned vector through an INDEPENDENT code path: the matrix-free element-wise
        #     K u (mfem_b200_apply_K = applyStiffnessMatrix, LinearElasticity.hh:801-823), free variables only
        if world == 1:
            Ku = np.asarray(h.apply_K(u.reshape(-1, 3))).reshape(-1)
            free = np.ones(u.size, dtype=bool); free[np.asarray(fixed)] = False
            fr = np.asarray(f_p).reshape(-1)
            parity["true_rel_residual_elementwise_apply_K"] = float(np.linalg.norm((fr - Ku)[free]) / np.linalg.norm(fr[free]))
        parity["min_uy"] = -rmax(-float(u.reshape(-1, 3)[:, 1].min()))
    h_keep = h if world > 1 else None          # N ranks: the e2e handles borrow this one's NCCL communicator
    if world == 1:
        h.close()

    # ------------------------------------------------------------------ end-to-end steps (host buffers)
    def e2e_step():
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        hh = new_handle(parent=h_keep)
        t1 = time.perf_counter()
        hh.assemble()
        t2 = time.perf_counter()
        hh.fix_variables(fixed, vals)
        uu, info = hh.solve(f_p, rtol=RTOL, return_info=True, out=u_p)
        t3 = time.perf_counter()
        tip = float(uu.reshape(-1, 3)[:, 1].min())
        parts = {"create_upload_reorder_s": t1 - t0, "symbolic_pattern_s": hh.timer("Pattern"),
                 "assemble_call_s": t2 - t1, "numeric_assembly_s": hh.timer("Assemble System"),
                 "preconditioner_setup_s": max(0.0, hh.timer("Fix Variables")) + max(0.0, hh.timer("Coarse Space")),
                 "coarse_structure_s": max(0.0, hh.timer("Coarse Structure")), "coarse_matrix_s": max(0.0, hh.timer("Coarse Matrix")),
                 "coarse_inverse_s": max(0.0, hh.timer("Coarse Inverse")), "coarse_level1_s": max(0.0, hh.timer("Coarse Level 1")),
                 "pcg_s": info[0]["seconds"], "fix_solve_download_call_s": t3 - t2}
        hh.close()
        if dist is not None:
            torch.cuda.synchronize(); dist.barrier()
        return time.perf_counter() - t0, tip, parts
    if args.warmup > 0:
        e2e_step()
    n_e2e = 2 if args.steps >= 2 else 1
    e2e_s, tip, parts = 0.0, None, None
    for _ in range(n_e2e):
        s, tip, parts = e2e_step()
        e2e_s += s
    e2e_max = rmax(e2e_s)
    tip_min = -rmax(-tip)
    h2d = rsum(nodes_p.nbytes + elems_p.nbytes + f_p.nbytes + np.asarray(fixed).nbytes + np.asarray(vals).nbytes)
    d2h = rsum(f_p.nbytes)
    parts = {k: rmax(v) for k, v in parts.items()}
    if h_keep is not None:
        h_keep.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_sample((60, 12, 12) if deg == 2 else (50, 10, 10), deg, mat, RTOL)
        cpu["pcg"] = {"elements_per_s": cpu["value"], "assembly_seconds": cpu["assembly_seconds"], "solve_seconds": cpu["solve_seconds"],
                      "iterations": cpu["pcg_iterations"]}
        for k in ("seconds", "elements", "pcg_iterations", "assembly_seconds", "solve_seconds"):
            cpu.pop(k, None)
        if not args.no_direct:
            cpu["direct"] = cpu_direct_sample(deg, mat)

    if rank == 0:
        spmv_bytes = nnzb_tot * 76 + nb_tot * 52           # whole-job algorithmic bytes of one (distributed) SpMV
        asm_bytes = nnzb_tot * 72 + n_elems * (4 * npe + 96)
        peak_all = hbm_peak * world
        rank_note = ", per-rank local part, max over ranks)" if world > 1 else ")"
        spmv_roof = {"kernel": "k_bsr_spmv<3,32,masked,dot> (block-CSR SpMV on the stored matrix" + rank_note,
                     "achieved": spmv_bytes / spmv_max / 1e9, "unit": "GB/s", "frac": spmv_bytes / spmv_max / 1e9 / peak_all,
                     "traffic": measured_traffic(name, "k_bsr_spmv", nnzb) if world == 1 else None,
                     "algorithmic_bytes_per_launch": spmv_bytes, "seconds_per_launch": spmv_max}
        asm_roof = {"achieved": asm_bytes / (asm_max / args.steps) / 1e9, "unit": "GB/s",
                    "frac": asm_bytes / (asm_max / args.steps) / 1e9 / peak_all, "algorithmic_bytes_per_launch": asm_bytes}
        if mf_used:
            # the iteration multiplies with the mesh-based operator: k_mf_elements (one thread per element: DoF ids 4*npe,
            # packed geometry 128, the element's npe result blocks 24*npe; the x blocks once per DoF) then k_mf_gather
            # (per incidence one 4-byte slot id + 24 bytes; per DoF row: extent 8, x 24, y 24, mask 3)
            n_inc = n_elems * npe
            if n_partials > 0:
                # chunked: per element the packed geometry (128) and two 16-bit tables per slot (4*npe); per (chunk, DoF)
                # partial its extent (2), DoF id (4) and result (24); the x blocks once per DoF
                elem_bytes = n_elems * (128 + 4 * npe) + n_partials * 30 + nb_tot * 24
                gather_bytes = n_partials * 28 + nb_tot * 59
                elem_kernel = "k_mf_chunk<3,2> (element kernel of the PCG's mesh-based operator: one CTA per 64 elements, per-chunk partial sums"
            else:
                elem_bytes = n_elems * (4 * npe + 128 + 24 * npe) + nb_tot * 24
                gather_bytes = n_inc * 28 + nb_tot * 59
                elem_kernel = "k_mf_elements<3,2> (element kernel of the PCG's mesh-based operator"
            roofline = {"bound": "hbm", "kernel": elem_kernel + rank_note,
                        "achieved": elem_bytes / op_elem_max / 1e9, "peak": peak_all, "unit": "GB/s",
                        "frac": elem_bytes / op_elem_max / 1e9 / peak_all,
                        "traffic": measured_traffic(name, "k_mf_chunk" if n_partials > 0 else "k_mf_elements", nnzb) if world == 1 else None,
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": elem_bytes, "seconds_per_launch": op_elem_max,
                        "gather_kernel": {"kernel": "k_mf_gather<3,masked,dot>", "achieved": gather_bytes / op_gather_max / 1e9, "unit": "GB/s",
                                          "frac": gather_bytes / op_gather_max / 1e9 / peak_all,
                                          "traffic": measured_traffic(name, "k_mf_gather", nnzb) if world == 1 else None,
                                          "algorithmic_bytes_per_launch": gather_bytes, "seconds_per_launch": op_gather_max},
                        "operator_seconds_per_product": op_max, "partials": int(n_partials), "operator_plan_s_once_per_mesh": mf_plan_s,
                        # the element kernel is not HBM-bound (ncu, profiles/r2_ncu_full_cfg5_matrix_free_pcg_kernels.txt: DRAM 30 %, L1 /
                        # shared memory 82.6 %, FP64 pipe 44.5 %): its FP64 rate beside the byte rate.  769 FP64 instructions per
                        # element in the SASS of k_mf_chunk<3,2> (546 DFMA + 126 DADD + 97 DMUL) = 1315 flop
                        "fp64": {"flop_per_launch": 1315 * n_elems, "achieved_tflops": 1315 * n_elems / op_elem_max / 1e12,
                                 "peak_tflops": 37.0 * world, "peak_source": "NVIDIA spec (FP64 vector), not in MEASURED_PEAKS.json",
                                 "frac": 1315 * n_elems / op_elem_max / 1e12 / (37.0 * world)},
                        "note": "the PCG multiplies with the mesh-based operator, so the dominant kernel of the step is its element kernel, "
                                "not the SpMV; SURVEY 8(d)'s SpMV figure (nnzb*76 + nb*52 bytes) is reported under stored_matrix_spmv, and "
                                "the same bytes over the operator's time under equivalent_stored_matrix_bandwidth",
                        "equivalent_stored_matrix_bandwidth": {"achieved": spmv_bytes / op_max / 1e9, "unit": "GB/s",
                                                               "frac": spmv_bytes / op_max / 1e9 / peak_all,
                                                               "note": "bytes the stored-matrix SpMV would have to stream for the same product / operator time"},
                        "stored_matrix_spmv": spmv_roof, "assembly": asm_roof}
        else:
            roofline = dict(spmv_roof)
            roofline.update({"bound": "hbm", "peak": peak_all, "peak_source": peak_src, "assembly": asm_roof})
            roofline["kernel"] = roofline["kernel"].replace("block-CSR SpMV on the stored matrix", "the PCG's SpMV")
        cfg = {"workload": workload_name(name, grid, deg, mat), "elements": n_elems, "nodes": n_nodes, "dofs": 3 * n_nodes,
               "nnz_blocks": int(nnzb_tot), "rtol": RTOL, "preconditioner": precond_name(args.coarse_aggregates, args.coarse_fine_nodes),
               "coarse_space": coarse_sizes,
               "pcg_operator": ("mesh-based (matrix-free): element kernel + per-DoF gather, csrc/matfree.inl" if mf_used
                                else "stored block-CSR matrix (SpMV)"), "l2_policy": "inputs larger than L2 (matrix %.1f GB)" % (nnzb_tot * 76 / 1e9 / world)}
        if world > 1:
            cfg["partition"] = (f"{world} x-slabs of elements, shared interface DoFs; per iteration one interface sum-exchange, two all-reduces "
                                f"and one all-gather, " + ("by the library's own kernels over NVLink peer memory (CUDA IPC window)" if uses_window
                                                           else "by NCCL"))
        out = {
            "metric": METRIC, "value": args.steps * n_elems / dev_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "assembly_elements_per_s": args.steps * n_elems / asm_max,
            "pcg_iters_per_s": iters / solve_max, "pcg_iterations_per_solve": iters / args.steps,
            "pcg_rel_residual": relres, "assembly_ms": 1e3 * asm_max / args.steps,
            "preconditioner_setup_ms": 1e3 * setup_max / args.steps, "solve_ms": 1e3 * solve_max / args.steps,
            "symbolic_pattern_ms": 1e3 * pattern_s, "wall_ms_per_step": 1e3 * wall / args.steps, "mesh_generation_s": t_gen,
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": n_e2e * n_elems / e2e_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "seconds_per_step": e2e_max / n_e2e, "steps": n_e2e, "min_uy": tip_min, "breakdown_last_step": parts,
                    "includes": "handle creation, mesh upload, DoF reordering, symbolic pattern, assembly, constraints, "
                                "preconditioner set-up, PCG, result download"},
            "parity": parity, "gpu_launches": int(launches_tot), "clocks": clocks,
        }
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg5")
    ap.add_argument("--matrix-free", type=int, default=-1,
                    help="PCG operator: -1 automatic (mesh-based for 3D quadratic elements), 0 stored-matrix SpMV, 1 mesh-based")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-direct", action="store_true", help="skip the direct-solver leg of the CPU baseline")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-size parity evidence (two extra solves)")
    ap.add_argument("--no-bj-parity", action="store_true", help="skip the block-Jacobi cross-check (48 s on one GPU for cfg5)")
    ap.add_argument("--coarse-aggregates", type=int, default=2048,
                    help="large aggregates of the multilevel preconditioner (0 = block-Jacobi only, -1 = automatic)")
    ap.add_argument("--coarse-fine-nodes", type=int, default=64, help="nodes per small (level-1) aggregate, 0 = none")
    ap.add_argument("--comm-p2p", type=int, default=1, help="N GPUs: 1 = small collectives by the library's kernels over peer memory, 0 = NCCL")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
