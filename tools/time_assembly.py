"""Assembly timing on the GPU box: python tools/time_assembly.py <cfg> [mode ...]
Prints per-mode numeric assembly time (median of 5, CUDA events of the library), the symbolic
(pattern + plan) time, and the algorithmic HBM fraction."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import meshfem_b200
import workloads as wl

cfg = sys.argv[1]
modes = [int(x) for x in sys.argv[2:]] or [0, 2]
grid, deg, mat = wl.CONFIGS[cfg]
m = wl.grid_femmesh(grid, deg)
D = wl.material(mat)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
ref = None
for mode in modes:
    h = meshfem_b200.Handle(0, assembly=mode)
    h.set_mesh(3, deg, m.nodes, m.elem_nodes)
    h.set_material(D)
    h.assemble()
    pat = h.timer("Pattern"); pad = h.timer("Plan Padding Ratio")
    ts = []
    for _ in range(5):
        h.reset_timers(); h.assemble(); ts.append(h.timer("Assemble System"))
    nb, nnzb = h.bsr_sizes()
    vals = h.get_bsr()[2] if m.num_elements <= 3_000_000 else None
    if vals is not None:
        if ref is None: ref = vals
        diff = float(np.abs(vals - ref).max() / np.abs(ref).max())
    else:
        diff = None
    t = float(np.median(ts))
    bytes_alg = nnzb * 72 + m.num_elements * (4 * m.elem_nodes.shape[1] + 96)
    print(json.dumps(dict(cfg=cfg, mode=mode, assemble_ms=round(t * 1e3, 3), pattern_ms=round(pat * 1e3, 1),
                          elems_per_s=round(m.num_elements / t), GBs=round(bytes_alg / t / 1e9, 1),
                          frac=round(bytes_alg / t / 1e9 / peak, 4), nnzb=nnzb, plan_padding=round(pad, 3), max_rel_diff_vs_first=diff)), flush=True)
    h.close()
