"""Scratch exploration on the GPU box: tolerance calibration and phase timings."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import meshfem_b200
import workloads as wl

def run(grid, deg, mat, rtols, direct=False, reorder=1):
    t = time.time(); m = wl.grid_femmesh(grid, deg); tm = time.time() - t
    D = wl.material(mat)
    fixed, vals, f = wl.cantilever_inputs(m)
    out = dict(grid=grid, deg=deg, elems=m.num_elements, nodes=m.num_nodes, mesh_s=round(tm, 2))
    h = meshfem_b200.Handle(0, reorder=reorder)
    t = time.time(); h.set_mesh(3, deg, m.nodes, m.elem_nodes); out["set_mesh_s"] = round(time.time() - t, 3)
    h.set_material(D)
    t = time.time(); h.assemble(); out["first_assemble_wall_s"] = round(time.time() - t, 3)
    out["pattern_s"] = h.timer("Pattern"); out["assemble_s"] = h.timer("Assemble System")
    h.reset_timers(); h.assemble(); out["assemble2_s"] = h.timer("Assemble System")
    nb, nnzb = h.bsr_sizes(); out["nnzb"] = nnzb
    h.fix_variables(fixed, vals)
    for kern, lanes in ((1, 32), (1, 16), (2, 0)):
        h.set_option("spmv_kernel", kern); h.set_option("spmv_lanes", lanes)
        spmv = h.time_spmv(20)
        out[f"spmv_ms_k{kern}l{lanes}"] = round(spmv * 1e3, 4)
        out[f"spmv_GBs_k{kern}l{lanes}"] = round((nnzb * 76 + nb * 52) / spmv / 1e9, 1)
    h.set_option("spmv_kernel", int(os.environ.get("SPMV_KERNEL", "0")))
    uref = None
    if direct:
        import meshfem_oracle as orc
        V, T = orc.grid_simplices(list(grid))
        sim = orc.Simulator(3, deg, V, T); sim.set_material(D)
        t = time.time(); K = sim.stiffness(); uref = orc.solve_fixed(K, f.reshape(-1), fixed, vals); out["direct_s"] = round(time.time() - t, 2)
    h.set_option("spmv_lanes", int(os.environ.get("SPMV_LANES", "0")))
    for rtol in rtols:
        u, info = h.solve(f, rtol=rtol, return_info=True)
        rec = dict(rtol=rtol, iters=info[0]["iterations"], solve_s=round(info[0]["seconds"], 4), relres=info[0]["rel_residual"])
        if uref is not None:
            rec["rel_l2_vs_direct"] = float(np.linalg.norm(u - uref) / np.linalg.norm(uref))
        Ku = h.spmv(u.reshape(-1, 3)).reshape(-1)
        free = np.ones(Ku.size, bool); free[fixed] = False
        rec["true_relres"] = float(np.linalg.norm((Ku - f.reshape(-1))[free]) / np.linalg.norm(f.reshape(-1)[free]))
        out.setdefault("solves", []).append(rec)
    h.close()
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    which = sys.argv[1]
    if which == "calib":
        run((40, 8, 8), 2, "iso", [1e-6, 1e-8, 1e-10, 1e-12], direct=True)
        run((40, 8, 8), 1, "iso", [1e-6, 1e-8, 1e-10], direct=True)
    elif which == "cfg2":
        run((100, 20, 20), 1, "iso", [1e-8])
    elif which == "cfg3":
        run((130, 26, 26), 2, "ortho", [1e-8])
    elif which == "cfg3nr":
        run((130, 26, 26), 2, "ortho", [1e-8], reorder=0)
    elif which == "cfg5":
        run((220, 44, 44), 2, "iso", [1e-8])
