"""CPU dry run of bench.py's own arm with a MOCK Handle (no GPU, no library call): exercises the control flow and the
assembly of the JSON line -- roofline object for both PCG operators, parity block, e2e block -- so that an edit to bench.py
cannot break the round-end bench unnoticed.  Run by tests/test_bench_dryrun.py in a subprocess."""
import json, os, sys, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import meshfem_b200

class MockHandle:
    def __init__(self, dev=0, **opts):
        self.opts = dict(opts); self.t = {}
    def set_option(self, k, v): self.opts[k] = v
    def set_mesh(self, N, deg, nodes, elems, **kw): self.N, self.nn, self.ne, self.npe = N, nodes.shape[0], elems.shape[0], elems.shape[1]
    def set_material(self, D): pass
    def assemble(self): self.t["Assemble System"] = 0.002; self.t["Pattern"] = 0.01
    def fix_variables(self, f, v): pass
    def bsr_sizes(self): return self.nn, 27 * self.nn
    def timer(self, name): return self.t.get(name, -1.0)
    def reset_timers(self): self.t = {}
    def launch_count(self): return 100
    def solve(self, f, rtol=1e-8, max_iters=1000, return_info=False, out=None):
        u = np.ones(self.nn * 3) if out is None else out
        u[...] = 1.0
        self.t["Fix Variables"] = 0.001; self.t["Coarse Space"] = 0.003
        info = [dict(iterations=10, converged=True, rel_residual=1e-9, seconds=0.05)]
        return (u, info) if return_info else u
    def time_spmv(self, n): return 1e-3
    def time_operator(self, n):
        mf = self.opts.get("matrix_free", -1) != 0
        if mf: self.t["Matrix-free Partials"] = 2.5 * self.nn
        return (5e-4, [3e-4, 2e-4], True) if mf else (1e-3, [0.0, 0.0], False)
    def coarse_array(self, name): return np.array([100.0, 10.0, 10.0])
    def apply_K(self, u): return np.zeros_like(u)
    def comm_uses_peer_window(self): return False
    def close(self): pass
    def __enter__(self): return self
    def __exit__(self, *a): pass

meshfem_b200.Handle = MockHandle
from meshfem_b200 import build as mb
mb.build_all = lambda: None
import bench
import argparse
for mf in (-1, 0):
    sys.argv = ["bench.py", "--config", "6x2x2:2:iso", "--steps", "2", "--warmup", "1", "--no-cpu-baseline", "--matrix-free", str(mf)]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    line = buf.getvalue().strip().splitlines()[-1]
    d = json.loads(line)
    print(mf, d["roofline"]["kernel"][:60], round(d["roofline"]["frac"], 4), sorted(d["roofline"].keys()))
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["parity"] is not None
print("DRYRUN_OK")
