"""N>1 host-side logic on CPU: element partition + interface lists (C++ Partition.hh) and a gloo
world-size-2/3 emulation of the distributed PCG (tests/mrank_cpu_worker.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("nparts", [2, 3, 4, 8])
def test_partition_invariants(lib_built, nparts):
    import workloads as wl
    from meshfem_b200 import hostlib
    m = wl.grid_femmesh((16, 2, 2), 2)
    parts = [hostlib.partition(m, nparts, r) for r in range(nparts)]
    # elements: a partition; nodes: owned exactly once
    allel = np.concatenate([p.elems for p in parts])
    assert np.array_equal(np.sort(allel), np.arange(m.num_elements))
    sizes = [p.num_elements for p in parts]
    assert max(sizes) - min(sizes) <= 1
    owned_ids = np.concatenate([p.nodes_global[p.owned.astype(bool)] for p in parts])
    assert np.array_equal(np.sort(owned_ids), np.arange(m.num_nodes))
    for r, p in enumerate(parts):
        # local connectivity refers to the same global nodes
        assert np.array_equal(p.nodes_global[p.elem_nodes], m.elem_nodes[p.elems])
        assert np.array_equal(p.nodes, m.nodes[p.nodes_global])
        for q, idx in p.shared.items():
            other = parts[q]
            assert np.array_equal(p.nodes_global[idx], other.nodes_global[other.shared[r]])   # same order on both sides
            assert np.all(np.diff(p.nodes_global[idx]) > 0)
        # owner = lowest sharing rank
        for q, idx in p.shared.items():
            if q < r:
                assert not p.owned[idx].any()


@pytest.mark.parametrize("nparts", [2, 3, 5, 8])
def test_rcb_partition_invariants_and_interface_size(lib_built, nparts):
    """Recursive coordinate bisection (Partition.hh rcbPartition, opt-in): same invariants as the slabs, balanced parts,
    ranks with more than two neighbours on a cube, and less interface than slabs there."""
    import workloads as wl
    from meshfem_b200 import hostlib
    m = wl.grid_femmesh((6, 6, 6), 2)
    parts = [hostlib.partition(m, nparts, r, method="rcb") for r in range(nparts)]
    slabs = [hostlib.partition(m, nparts, r, method="slab") for r in range(nparts)]
    assert np.array_equal(np.sort(np.concatenate([p.elems for p in parts])), np.arange(m.num_elements))
    sizes = [p.num_elements for p in parts]
    assert max(sizes) - min(sizes) <= nparts                                   # floor/ceil splits at every level
    owned_ids = np.concatenate([p.nodes_global[p.owned.astype(bool)] for p in parts])
    assert np.array_equal(np.sort(owned_ids), np.arange(m.num_nodes))
    for r, p in enumerate(parts):
        assert np.array_equal(p.nodes_global[p.elem_nodes], m.elem_nodes[p.elems])
        for q, idx in p.shared.items():
            other = parts[q]
            assert r in other.shared
            assert np.array_equal(p.nodes_global[idx], other.nodes_global[other.shared[r]])
            if q < r:
                assert not p.owned[idx].any()
    if nparts == 8:
        assert max(len(p.shared) for p in parts) > 2                           # 2x2x2 boxes: every box touches the 7 others
        worst = lambda ps: max(sum(len(i) for i in p.shared.values()) for p in ps)
        assert worst(parts) < worst(slabs) * 1.01                              # per-rank exchange volume no larger than a slab's two faces
        distinct = lambda ps: max(len(np.unique(np.concatenate(list(p.shared.values())))) for p in ps)
        assert distinct(parts) < distinct(slabs)                               # fewer interface DoFs per rank
    # determinism
    again = hostlib.partition(m, nparts, 0, method="rcb")
    assert np.array_equal(again.elems, parts[0].elems)


@pytest.mark.parametrize("nparts", [2, 3, 4])
@pytest.mark.parametrize("deg", [1, 2])
def test_periodic_partition_works_on_dofs(lib_built, nparts, deg):
    """Periodic cell: identified nodes are one DoF BEFORE partitioning -- the x-slab wrap
    makes the first and the last rank neighbours and cell edges/corners are shared by several ranks."""
    from meshfem_b200 import hostlib
    raw = hostlib.grid([4, 4, 4], (0, 0, 0), (1, 1, 1))
    info = raw.apply_bc(deg, "", periodic=True)
    m, dfn, nd = info["mesh"], info["dof_for_node"], info["num_dofs"]
    assert nd < m.num_nodes
    parts = [hostlib.partition(m, nparts, r, dof_for_node=dfn) for r in range(nparts)]
    owned_ids = np.concatenate([p.dofs_global[p.owned.astype(bool)] for p in parts])
    assert np.array_equal(np.sort(owned_ids), np.arange(nd))                 # every DoF owned exactly once
    assert np.array_equal(np.sort(np.concatenate([p.elems for p in parts])), np.arange(m.num_elements))
    for r, p in enumerate(parts):
        assert p.num_dofs == p.dofs_global.size and np.all(np.diff(p.dofs_global) > 0)
        assert np.array_equal(p.dofs_global[p.dof_for_node], dfn[p.nodes_global])   # local map == global map
        assert np.array_equal(p.nodes_global[p.elem_nodes], m.elem_nodes[p.elems])
        for q, idx in p.shared.items():
            other = parts[q]
            assert np.array_equal(p.dofs_global[idx], other.dofs_global[other.shared[r]])
            if q < r:
                assert not p.owned[idx].any()
    if nparts > 2:     # the wrap: rank 0 and the last rank share the identified x-faces
        assert (nparts - 1) in parts[0].shared and 0 in parts[-1].shared


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_pcg_gloo(lib_built, world):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "mrank_cpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "MRANK_CPU" in r.stdout


def test_distributed_pcg_gloo_rcb_many_neighbours(lib_built):
    """The same emulation on 4 ranks of a cube cut by recursive coordinate bisection: ranks with three neighbours, DoFs
    shared by up to four ranks (exchange-add and owner mask must still count every DoF once)."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1", MESHFEM_PARTITIONER="rcb", MRANK_GRID="4,4,4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "mrank_cpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "MRANK_CPU" in r.stdout


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_periodic_homogenization_gloo(lib_built, world):
    """tests/mrank_cpu_homog_worker.py: DoF-based partition of a periodic cell, pinned variable, exchanged
    constant-strain loads, distributed PCG and the all-reduced volume form against the golden tensor."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29620 + world), os.path.join(ROOT, "tests", "mrank_cpu_homog_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "MRANK_CPU_HOMOG" in r.stdout
