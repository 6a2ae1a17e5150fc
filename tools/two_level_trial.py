#!/usr/bin/env python
"""Trial of the optional two-level preconditioner (option coarse_aggregates, csrc/coarse.inl) on a bench workload,
run by bench.py in a SUBPROCESS with a timeout so that a failure of this not-yet-default path cannot touch the
headline measurement.  Prints one JSON line: either the measurement or {"error": ...}.

  python tools/two_level_trial.py --config cfg5 --aggregates 2048 [--expect-min-uy V] [--device 0]

Protocol: one warm-up step (builds the symbolic pattern and the coarse space), then `--steps` timed steps of
numeric assembly + PCG solve with all inputs resident in HBM (the same step bench.py times for `value`), then one
end-to-end pass on a fresh handle from host buffers (the same pass bench.py times for `e2e`, the coarse-space setup
included).  The solution is validated by the tip deflection against the block-Jacobi solve of the parent
(--expect-min-uy, relative 1e-6 = the north-star displacement tolerance) and by the TRUE residual
||f_f - (K u)_f|| / ||f_f|| recomputed through the assembled SpMV from the downloaded solution.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg5")
    ap.add_argument("--aggregates", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--expect-min-uy", type=float, default=None)
    ap.add_argument("--shape", type=int, default=0, help="0 = near-cubic boxes (default), 1 = contiguous runs of the internal numbering")
    args = ap.parse_args()

    import meshfem_b200
    import workloads as wl
    from bench import parse_config

    name, grid, deg, mat = parse_config(args.config)
    m = wl.grid_femmesh(grid, deg)
    D = wl.material(mat)
    fixed, vals, f = wl.cantilever_inputs(m)
    n_elems = m.num_elements
    out = {"config": name, "aggregates": args.aggregates, "aggregate_shape": "boxes" if args.shape == 0 else "runs", "rtol": args.rtol,
           "elements": int(n_elems)}

    with meshfem_b200.Handle(args.device) as h:
        h.set_mesh(3, deg, m.nodes, m.elem_nodes)
        h.set_material(D)
        h.assemble()
        h.fix_variables(fixed, vals)
        h.set_option("coarse_shape", args.shape)
        h.set_option("coarse_aggregates", args.aggregates)
        h.reset_timers()
        h.solve(f, rtol=args.rtol, max_iters=8000)                                  # warm-up: builds the coarse space
        out["coarse_setup_ms"] = 1e3 * h.timer("Coarse Space")
        asm_s = solve_s = coarse_s = 0.0
        iters = 0
        for _ in range(args.steps):
            h.reset_timers()
            h.assemble()
            u, info = h.solve(f, rtol=args.rtol, max_iters=8000, return_info=True)
            asm_s += h.timer("Assemble System"); solve_s += info[0]["seconds"]; iters += info[0]["iterations"]
            coarse_s += h.timer("Coarse Space")      # re-assembly invalidates E = Z'KZ: rebuilt inside every step, and counted
        # true residual on the free variables through the assembled matrix
        Ku = np.asarray(h.spmv(u)).reshape(-1)
        free = np.ones(Ku.size, dtype=bool); free[np.asarray(fixed)] = False
        fr = np.asarray(f).reshape(-1)
        true_res = float(np.linalg.norm((fr - Ku)[free]) / np.linalg.norm(fr[free]))
        tip = float(np.asarray(u).reshape(-1, 3)[:, 1].min())
    out.update({
        "steps": args.steps, "pcg_iterations_per_solve": iters / args.steps, "solve_ms": 1e3 * solve_s / args.steps,
        "assembly_ms": 1e3 * asm_s / args.steps, "coarse_setup_ms_in_step": 1e3 * coarse_s / args.steps,
        "value": args.steps * n_elems / (asm_s + coarse_s + solve_s), "unit": "elements/s",
        "pcg_rel_residual": info[0]["rel_residual"], "true_rel_residual": true_res, "min_uy": tip,
    })

    t0 = time.perf_counter()
    with meshfem_b200.Handle(args.device, coarse_shape=args.shape, coarse_aggregates=args.aggregates) as hh:
        hh.set_mesh(3, deg, m.nodes, m.elem_nodes)
        hh.set_material(D)
        hh.assemble()
        hh.fix_variables(fixed, vals)
        u2 = hh.solve(f, rtol=args.rtol, max_iters=8000)
        tip2 = float(np.asarray(u2).reshape(-1, 3)[:, 1].min())
    e2e_s = time.perf_counter() - t0
    out["e2e"] = {"value": n_elems / e2e_s, "unit": "elements/s", "seconds_per_step": e2e_s, "min_uy": tip2,
                  "includes": "handle creation, mesh upload, DoF reordering, symbolic pattern, assembly, constraints, "
                              "coarse-space setup, PCG, result download"}
    ok = true_res <= 10 * args.rtol and abs(tip2 - tip) <= 1e-6 * abs(tip)
    if args.expect_min_uy is not None:
        out["min_uy_rel_diff_vs_block_jacobi"] = abs(tip - args.expect_min_uy) / abs(args.expect_min_uy)
        ok = ok and out["min_uy_rel_diff_vs_block_jacobi"] <= 1e-6
    out["valid"] = bool(ok)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    try:
        main()
    except BaseException as e:  # noqa: BLE001  (report, never raise: the parent only reads the JSON line)
        print(json.dumps({"error": f"{type(e).__name__}: {e}"[:400]}), flush=True)
        sys.exit(0)
