"""Batched (SpMM) vs one-at-a-time PCG on the GPU box: python tools/time_batched.py <cfg> [rtol]
Six right-hand sides against the cantilever matrix of <cfg>; prints device seconds of both modes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import meshfem_b200
import workloads as wl

cfg = sys.argv[1]
rtol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-8
grid, deg, mat = wl.CONFIGS[cfg]
m = wl.grid_femmesh(grid, deg)
fixed, vals, f = wl.cantilever_inputs(m)
rng = np.random.default_rng(0)
dirs = rng.normal(size=(6, 3))
rhs = np.stack([f * 0 + np.where(np.abs(f).sum(axis=1, keepdims=True) > 0, 1.0, 0.0) * d[None, :] for d in dirs])   # six tip loads
res = {}
for batch, kern in ((1, 0), (0, 0)):
    h = meshfem_b200.Handle(0, batch_rhs=batch, spmm_kernel=kern)
    h.set_mesh(3, deg, m.nodes, m.elem_nodes)
    h.set_material(wl.material(mat))
    h.assemble()
    h.fix_variables(fixed, vals)
    u, info = h.solve(rhs, rtol=rtol, return_info=True)
    res[batch] = (u, info)
    print(json.dumps(dict(cfg=cfg, batch=batch, spmm_kernel=kern, seconds=round(sum(i["seconds"] for i in info), 3),
                          iterations=[i["iterations"] for i in info], relres=[float("%.2e" % i["rel_residual"]) for i in info])), flush=True)
    h.close()
d = max(np.linalg.norm(res[1][0][k] - res[0][0][k]) / np.linalg.norm(res[0][0][k]) for k in range(6))
print(json.dumps(dict(max_rel_diff_batched_vs_sequential=float(d))))
