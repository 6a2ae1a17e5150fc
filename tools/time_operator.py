"""Stored-matrix SpMV vs the mesh-based (matrix-free) operator on the GPU box, plus one solve with each:
python tools/time_operator.py <cfg>"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import meshfem_b200
import workloads as wl

cfg = sys.argv[1]
grid, deg, mat = wl.CONFIGS[cfg]
m = wl.grid_femmesh(grid, deg)
fixed, vals, f = wl.cantilever_inputs(m)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
h = meshfem_b200.Handle(0, coarse_aggregates=-1 if cfg != "cfg5" else 2048)
h.set_mesh(3, deg, m.nodes, m.elem_nodes)
h.set_material(wl.material(mat))
h.assemble()
h.fix_variables(fixed, vals)
nb, nnzb = h.bsr_sizes()
ne, npe = m.num_elements, m.elem_nodes.shape[1]
t = min(h.time_spmv(20) for _ in range(2))
by = nnzb * 76 + nb * 52
print(json.dumps(dict(cfg=cfg, what="stored-matrix SpMV", ms=round(t * 1e3, 4), GBs=round(by / t / 1e9, 1), frac=round(by / t / 1e9 / peak, 4))), flush=True)
h.set_option("matrix_free", 1)
variants = [tuple(int(v) for v in a.split(":")) for a in sys.argv[2:]] or [(1, 0, 4, 0, 3, 128, 12), (1, 0, 4, 0, 3, 64, 12), (1, 0, 4, 0, 3, 32, 12), (1, 0, 4, 0, 3, 128, 16), (1, 0, 4, 0, 3, 64, 16), (1, 0, 4, 0, 3, 32, 16)]
best = None
for chunked, pad, lanes, order, pol, ch, cw in variants:
    h.set_option("mf_chunked", chunked); h.set_option("mf_chunk_elems", ch); h.set_option("mf_chunk_warps", cw); h.set_option("mf_slot_pad", pad); h.set_option("mf_gather_lanes", lanes); h.set_option("mf_elem_order", order); h.set_option("mf_gather_policy", pol)
    h.time_operator(3)
    r = [h.time_operator(20) for _ in range(2)]
    top = min(x[0] for x in r); te = min(x[1][0] for x in r); tg = min(x[1][1] for x in r)
    ss = 32 if pad else 24
    be = ne * (4 * npe + 128 + ss * npe) + nb * 24
    bg = ne * npe * (ss + 4) + nb * (8 + 24 + 24 + 3)
    if chunked:
        nP = int(h.timer("Matrix-free Partials"))
        be = ne * (128 + 4 * npe) + nP * (2 + 4 + 24) + nb * 24      # geometry, 2 x 16-bit tables per slot, per partial: extent, DoF id, result; x once
        bg = nP * (24 + 4) + nb * (8 + 24 + 24 + 3)
    print(json.dumps(dict(cfg=cfg, what="matrix-free operator", chunked=chunked, chunk_elems=ch, chunk_warps=cw, partials=(int(h.timer("Matrix-free Partials")) if chunked else 0), slot_pad=pad, gather_lanes=lanes, elem_order=order, slot_policy=pol, ms=round(top * 1e3, 4), elements_ms=round(te * 1e3, 4),
                          gather_ms=round(tg * 1e3, 4), elements_GBs=round(be / te / 1e9, 1), elements_frac=round(be / te / 1e9 / peak, 4),
                          gather_GBs=round(bg / tg / 1e9, 1), gather_frac=round(bg / tg / 1e9 / peak, 4), speedup=round(t / top, 3))), flush=True)
    if best is None or top < best[0]:
        best = (top, chunked, pad, lanes, order, pol, ch, cw)
h.set_option("mf_chunked", best[1]); h.set_option("mf_chunk_elems", best[6]); h.set_option("mf_chunk_warps", best[7]); h.set_option("mf_slot_pad", best[2]); h.set_option("mf_gather_lanes", best[3]); h.set_option("mf_elem_order", best[4]); h.set_option("mf_gather_policy", best[5])
print(json.dumps(dict(best=best, plan_s=h.timer("Matrix-free Plan"), partials=int(h.timer("Matrix-free Partials")))), flush=True)
us = {}
for mf in (0, 1):
    h.set_option("matrix_free", mf)
    for rep in range(2):
        u, info = h.solve(f, rtol=1e-8, return_info=True)
    us[mf] = np.asarray(u).reshape(-1)
    print(json.dumps(dict(cfg=cfg, matrix_free=mf, iterations=info[0]["iterations"], solve_s=round(info[0]["seconds"], 4),
                          ms_per_iteration=round(1e3 * info[0]["seconds"] / info[0]["iterations"], 4), rel_residual=info[0]["rel_residual"])), flush=True)
print(json.dumps(dict(cfg=cfg, rel_l2_matrix_free_vs_stored=float(np.linalg.norm(us[1] - us[0]) / np.linalg.norm(us[0])))), flush=True)
h.close()
