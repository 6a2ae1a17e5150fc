"""Mid-size golden displacement field for the parity-at-scale GPU tests: the oracle's DIRECT solve (scipy splu,
symmetric-mode MMD ordering; stand-in for CHOLMOD, the SPD solution is unique) of the quadratic-tet cantilever
`grid 40x8x8 -t` (61,440 elements, 92,785 nodes, 278,355 DoF), isotropic E = 200, nu = 0.35,
examples/cantilever/cantilever.bc.  Takes ~3 minutes and ~6 GB on one core, which is why the result is committed
(float64, compressed) instead of being recomputed by the tests.

  python tests/golden/make_midsize_golden.py
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from util import cantilever_problem

SIZES = (40, 8, 8)
t = time.time()
sim, fixed, vals, f = cantilever_problem(3, 2, SIZES)
u = sim.solve(f)
K = sim.stiffness()
free = np.ones(K.shape[0], bool); free[fixed] = False
res = float(np.linalg.norm((f.reshape(-1) - K @ u.reshape(-1))[free]) / np.linalg.norm(f.reshape(-1)[free]))
print(f"direct solve of {u.shape[0]} nodes in {time.time() - t:.0f} s, relative residual {res:.2e}, min u_y {u[:, 1].min():.12f}")
np.savez_compressed(os.path.join(HERE, "cantilever_40x8x8_deg2.npz"), sizes=np.array(SIZES), u=u, rel_residual=res)
