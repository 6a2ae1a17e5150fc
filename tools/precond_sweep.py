#!/usr/bin/env python
"""Sweep of the multilevel preconditioner's two knobs on a bench workload (one mesh build, one symbolic phase):

  python tools/precond_sweep.py --config cfg5 --combos 0:0,2048:0,2048:32,4096:32,1024:48

Each combo is `coarse_aggregates:coarse_fine_nodes` (0:0 = block-Jacobi only, -1:32 = automatic).  Per combo one
warm-up solve (builds the coarse space) and one timed step of numeric assembly + coarse set-up + PCG with the inputs
resident in HBM; prints one JSON line with the iteration count, the times of the step's parts (CUDA-event timers of
the library), the TRUE residual ||f_f - (K u)_f|| / ||f_f|| recomputed through the assembled SpMV and the full-vector
relative L2 distance to the first combo's solution.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402

COARSE_TIMERS = ("Coarse Space", "Coarse Structure", "Coarse Matrix", "Coarse Inverse", "Coarse Level 1", "Fix Variables")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg5")
    ap.add_argument("--combos", default="2048:0,2048:32")
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--max-iters", type=int, default=10000)
    args = ap.parse_args()

    import meshfem_b200
    import workloads as wl
    from bench import parse_config

    name, grid, deg, mat = parse_config(args.config)
    m = wl.grid_femmesh(grid, deg)
    D = wl.material(mat)
    fixed, vals, f = wl.cantilever_inputs(m)
    n_elems = m.num_elements
    free = np.ones(f.size, dtype=bool); free[np.asarray(fixed)] = False
    fr = np.asarray(f).reshape(-1)
    u_first = None
    with meshfem_b200.Handle(args.device) as h:
        h.set_mesh(3, deg, m.nodes, m.elem_nodes)
        h.set_material(D)
        h.assemble()
        h.fix_variables(fixed, vals)
        for combo in args.combos.split(","):
            S, fine = (int(x) for x in combo.split(":"))
            out = {"config": name, "aggregates": S, "fine_nodes": fine, "rtol": args.rtol, "elements": int(n_elems)}
            try:
                h.set_option("coarse_aggregates", S)
                h.set_option("coarse_fine_nodes", fine)
                h.reset_timers()
                h.solve(f, rtol=args.rtol, max_iters=args.max_iters)                 # warm-up: builds the coarse space
                out["first_setup_ms"] = {k: round(1e3 * h.timer(k), 3) for k in COARSE_TIMERS}
                h.reset_timers()
                h.assemble()
                u, info = h.solve(f, rtol=args.rtol, max_iters=args.max_iters, return_info=True)
                asm_s, solve_s = h.timer("Assemble System"), info[0]["seconds"]
                setup = {k: round(1e3 * h.timer(k), 3) for k in COARSE_TIMERS}
                coarse_s = h.timer("Coarse Space") + h.timer("Fix Variables")
                Ku = np.asarray(h.spmv(u)).reshape(-1)
                u = np.asarray(u).reshape(-1)
                if u_first is None:
                    u_first = u.copy()
                out.update({
                    "iterations": info[0]["iterations"], "solve_ms": round(1e3 * solve_s, 2), "assembly_ms": round(1e3 * asm_s, 3),
                    "setup_ms_in_step": setup, "step_ms": round(1e3 * (asm_s + coarse_s + solve_s), 2),
                    "ms_per_iteration": round(1e3 * solve_s / max(info[0]["iterations"], 1), 4),
                    "elements_per_s": n_elems / (asm_s + coarse_s + solve_s),
                    "pcg_rel_residual": info[0]["rel_residual"],
                    "true_rel_residual": float(np.linalg.norm((fr - Ku)[free]) / np.linalg.norm(fr[free])),
                    "rel_l2_vs_first": float(np.linalg.norm(u - u_first) / np.linalg.norm(u_first)),
                    "min_uy": float(u.reshape(-1, 3)[:, 1].min()),
                })
            except Exception as e:  # noqa: BLE001
                out["error"] = f"{type(e).__name__}: {e}"[:400]
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
