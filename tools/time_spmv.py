"""SpMV timing on the GPU box (the PCG's masked + fused-dot launch): python tools/time_spmv.py <cfg> [kernel:lanes[:prefetch] ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import meshfem_b200
import workloads as wl

cfg = sys.argv[1]
variants = [tuple(int(v) for v in a.split(":")) for a in sys.argv[2:]] or [(1, 32), (3, 32), (2, 0)]
grid, deg, mat = wl.CONFIGS[cfg]
m = wl.grid_femmesh(grid, deg)
fixed, vals, f = wl.cantilever_inputs(m)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
h = meshfem_b200.Handle(0)
h.set_mesh(3, deg, m.nodes, m.elem_nodes)
h.set_material(wl.material(mat))
h.assemble()
h.fix_variables(fixed, vals)
nb, nnzb = h.bsr_sizes()
by = nnzb * 76 + nb * 52
for var in variants:
    kern, lanes, pf, minb = (tuple(var) + (0, 0))[:4]
    h.set_option("spmv_kernel", kern); h.set_option("spmv_lanes", lanes); h.set_option("spmv_prefetch", pf); h.set_option("spmv_min_blocks", minb)
    h.time_spmv(5)
    t = min(h.time_spmv(20) for _ in range(3))
    print(json.dumps(dict(cfg=cfg, kernel=kern, lanes=lanes, prefetch=pf, min_blocks=minb, spmv_ms=round(t * 1e3, 4), GBs=round(by / t / 1e9, 1), frac=round(by / t / 1e9 / peak, 4))), flush=True)
h.close()
