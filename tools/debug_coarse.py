"""GPU debugging aid: the device's coarse-space arrays against the numpy emulation (tools/emulate_multilevel.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
import meshfem_b200
import emulate_multilevel as em
from util import cantilever_problem

for N, deg, sizes, S, fine in [(3, 2, (20, 4, 4), 16, 60), (2, 2, (40, 8), 96, 12)]:
    sim, fixed, vals, f = cantilever_problem(N, deg, sizes)
    op, info = em.full_operator(sim, fixed, S, fine)
    A = info["arrays"]
    M = 6 if N == 3 else 3
    with meshfem_b200.Handle(0, coarse_aggregates=S, coarse_fine_nodes=fine) as h:
        h.set_mesh(N, deg, sim.mesh.nodes, sim.mesh.elem_nodes)
        h.set_material(sim.D); h.assemble(); h.fix_variables(fixed, vals)
        rng = np.random.default_rng(5)
        r = rng.standard_normal(sim.mesh.num_nodes * N)
        z, rz = h.apply_preconditioner(r)
        sz = h.coarse_array("sizes"); print("case", N, deg, sizes, S, fine, "device sizes", sz, "emulation", info["S1"], info["S2"], info["b"], info["r"])
        i2e = h.coarse_array("int2ext").astype(np.int64)
        agg1 = h.coarse_array("agg1").astype(np.int64)
        Y1 = h.coarse_array("Y1").reshape(-1, N)
        slot_dev = np.empty_like(agg1); slot_dev[i2e] = agg1
        Y1_dev = np.empty_like(Y1); Y1_dev[i2e] = Y1
        print("  slots equal:", np.array_equal(slot_dev, A["slot"]), " mismatches", int((slot_dev != A["slot"]).sum()))
        print("  Y1 max diff:", np.abs(Y1_dev - A["Y1"]).max())
        sh = h.coarse_array("shift").reshape(-1, N)
        print("  shift max diff:", np.abs(sh[:A["shift"].shape[0]] - A["shift"]).max())
        Einv = h.coarse_array("Einv").reshape(A["Einv"].shape)
        print("  Einv rel diff:", np.abs(Einv - A["Einv"]).max() / np.abs(A["Einv"]).max())
        B = h.coarse_array("B1inv").reshape(-1, M, M)
        Be = A["B1inv"]
        d = np.abs(B - Be).reshape(B.shape[0], -1).max(axis=1)
        nrm = np.abs(Be).reshape(B.shape[0], -1).max(axis=1)
        bad = np.nonzero(d > 1e-8 * np.maximum(nrm, 1e-300))[0]
        print("  B1inv blocks:", B.shape[0], "bad:", bad.size, "worst rel", (d / np.maximum(nrm, 1e-300)).max())
        if bad.size:
            s = bad[0]
            np.set_printoptions(precision=4, linewidth=200)
            print("  slot", s, "device B1inv\n", B[s], "\n  emulation\n", Be[s])
            print("  device inverse^-1 (D1 as the device saw it)\n", np.linalg.pinv(B[s]), "\n  emulation D1\n", np.linalg.pinv(Be[s]))
        if N == 3:
            D = h.coarse_array("D1").reshape(-1, M * (M + 1) // 2)
            Dm = np.zeros((D.shape[0], M, M)); t = 0
            for b in range(M):
                for a in range(b + 1):
                    Dm[:, a, b] = Dm[:, b, a] = D[:, t]; t += 1
            dd = np.abs(Dm - A["D1"]).reshape(D.shape[0], -1).max(axis=1) / np.abs(A["D1"]).reshape(D.shape[0], -1).max(axis=1)
            print("  D1 bad blocks:", int((dd > 1e-9).sum()), "worst", dd.max())
            s = int(np.argmax(dd))
            print("  slot", s, "device D1\n", Dm[s], "\n emulation D1\n", A["D1"][s], "\n ratio\n", Dm[s] / A["D1"][s])
            print("  inv check: |B1inv_dev - pinv(D1_dev)| =", np.abs(B[s] - np.linalg.pinv(Dm[s])).max(), "eig(D1_dev)", np.linalg.eigvalsh(Dm[s]))
        e, erz = op(r)
        print("  operator rel err", np.linalg.norm(z.reshape(-1) - e) / np.linalg.norm(e), "rz", rz, erz)
