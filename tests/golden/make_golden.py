"""Generates the golden vectors under tests/golden/ from the CPU oracle (run in the build
container: `python tests/golden/make_golden.py`).  The reference itself cannot be built here
(no Eigen/SuiteSparse/TBB/Boost offline), so these are oracle outputs -- restatements of the
reference algorithm checked against its own KATs in tests/test_oracle_kats.py -- for the
configurations BASELINE.json names, at sizes a direct solver finishes in seconds:
  cfg1_deg{1,2}.npz   examples/cantilever/square.msh (== grid 10x10 -t) + cantilever_2D.bc + B9Creator
  cant3d_deg{1,2}.npz 3D cantilever grid 10x2x2 (the config-2/3/5 family), isotropic / orthotropic
  homog_perforated.npz periodic cell 4^3 hexes minus the centred 2^3 block, quadratic + linear tets
"""
import itertools
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))
import meshfem_oracle as orc  # noqa: E402
from util import CANTILEVER_2D_BC, CANTILEVER_BC, ORTHO  # noqa: E402


def perforated_cell(n=4):
    V, H = orc.gen_grid([n, n, n])
    lo, hi = n // 4, n - n // 4
    keep = [i for i, (s, r, c) in enumerate(itertools.product(range(n), range(n), range(n)))
            if not (lo <= s < hi and lo <= r < hi and lo <= c < hi)]
    Vt, T = orc.hex_tet_subdiv(V, H[keep])
    used = np.unique(T)
    remap = -np.ones(Vt.shape[0], dtype=np.int64); remap[used] = np.arange(used.size)
    return Vt[used] / n, remap[T]


def main():
    for deg in (1, 2):
        V, T = orc.grid_simplices([10, 10])
        r = orc.simulate(2, deg, V, T, orc.isotropic_D(2, 200.0, 0.35), CANTILEVER_2D_BC)
        np.savez_compressed(os.path.join(HERE, f"cfg1_deg{deg}.npz"), u=r["u"], load=r["load"], strain=r["strain"],
                            stress=r["stress"], Ku=r["Ku"])
        V, T = orc.grid_simplices([10, 2, 2])
        D = orc.isotropic_D(3, 200.0, 0.35) if deg == 1 else orc.material_from_json(3, ORTHO)
        r = orc.simulate(3, deg, V, T, D, CANTILEVER_BC)
        np.savez_compressed(os.path.join(HERE, f"cant3d_deg{deg}.npz"), u=r["u"], load=r["load"], strain=r["strain"],
                            stress=r["stress"], Ku=r["Ku"])
    out = {}
    V, T = perforated_cell(4)
    for deg in (1, 2):
        sim = orc.Simulator(3, deg, V, T)
        sim.set_material(orc.isotropic_D(3, 200.0, 0.35))
        w = orc.solve_cell_problems(sim)
        out[f"Eh_deg{deg}"] = orc.homogenized_tensor_displacement_form(sim, w)
        out[f"w_deg{deg}"] = np.stack(w)
    np.savez_compressed(os.path.join(HERE, "homog_perforated.npz"), V=V, T=T, **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
