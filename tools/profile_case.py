"""Small driver for ncu captures: python tools/profile_case.py <cfg> <what>
what = spmv | assemble | solve  (solve: 2 x 25 PCG iterations)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import meshfem_b200
import workloads as wl

cfg, what = sys.argv[1], sys.argv[2]
grid, deg, mat = wl.CONFIGS[cfg]
m = wl.grid_femmesh(grid, deg)
fixed, vals, f = wl.cantilever_inputs(m)
h = meshfem_b200.Handle(0)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    h.set_option(k, int(v))
h.set_mesh(3, deg, m.nodes, m.elem_nodes)
h.set_material(wl.material(mat))
h.assemble()
h.fix_variables(fixed, vals)
if what == "spmv":
    print("spmv s/launch", h.time_spmv(5))
elif what == "assemble":
    h.assemble()
elif what == "solve6":
    import numpy as np
    rhs = np.stack([f * (k + 1.0) for k in range(6)])
    try:
        h.solve(rhs, rtol=1e-30, max_iters=50)
    except meshfem_b200.MfemB200Error as e:
        print("expected:", e)
elif what == "solve":
    try:
        h.solve(f, rtol=1e-30, max_iters=50)
    except meshfem_b200.MfemB200Error as e:
        print("expected:", e)
h.close()
