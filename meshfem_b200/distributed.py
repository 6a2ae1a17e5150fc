"""Element-partitioned execution on several GPUs of one box: one process per GPU, torch.distributed
for rendezvous / barriers / host-side reductions, the library's own NCCL communicator for the data
path (interface sum-exchange inside the PCG, all-reduced dots; csrc/comm.cu).

* `local_problem` / `make_handle`: slice a global problem for this rank and build its handle
  (what bench.py and the multi-rank tests use for the cantilever configs).
* `homogenize`: periodic homogenization of a base cell (PeriodicHomogenization.hh:34-54 solveCellProblems
  + :72-100 homogenizedElasticityTensor, the volume form) on 1..N GPUs.  Periodically identified nodes
  are ONE DoF before partitioning (the partitioner works on DoFs), so the periodic wrap simply makes the
  first and last slab neighbours.
"""
from __future__ import annotations

import numpy as np

from . import hostlib
from .capi import Handle, flat_len


def broadcast_bytes(dist, payload, n, device=None):
    import torch
    t = torch.zeros(n, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def local_problem(m, fixed, vals, f, world, rank, dof_for_node=None):
    """Slice the global problem (FEMMesh data m, fixed VARIABLES/values in global DoF numbering,
    consistent per-DoF load f) for `rank`.  Returns (partition, local fixed vars, values, local f)."""
    p = hostlib.partition(m, world, rank, dof_for_node=dof_for_node)
    N = m.N
    fixed = np.asarray(fixed, dtype=np.int64)
    fdof, fcomp = fixed // N, fixed % N
    pos = np.searchsorted(p.dofs_global, fdof)
    pos = np.minimum(pos, p.num_dofs - 1)
    keep = p.dofs_global[pos] == fdof
    lfixed = (N * pos[keep] + fcomp[keep]).astype(np.int64)
    lvals = np.asarray(vals, dtype=np.float64)[keep]
    lf = None if f is None else np.ascontiguousarray(np.asarray(f).reshape(-1, N)[p.dofs_global])
    return p, lfixed, lvals, lf


def make_handle(dist, world, rank, local_rank, p, D, comm_parent=None, **options):
    """Handle with communicator, local mesh, interface and material.  comm_parent: a live handle of this
    process whose NCCL communicator is reused (no second ncclCommInitRank)."""
    import torch
    h = Handle(local_rank, **options)
    if comm_parent is not None:
        h.comm_share(comm_parent)
    else:
        uid = Handle.comm_unique_id() if rank == 0 else None
        dev = torch.device("cuda", local_rank) if dist.get_backend() == "nccl" else None
        uid = broadcast_bytes(dist, uid, 128, dev)
        h.comm_init(world, rank, uid)
        if dev is not None and world <= 8 and options.get("comm_p2p", 1):
            # peer window: the IPC handles of all ranks, gathered in rank order (csrc/comm.cu)
            mine = torch.frombuffer(bytearray(h.comm_window_handle()), dtype=torch.uint8).to(dev)
            parts = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(parts, mine)
            h.comm_window_open(b"".join(bytes(t.cpu().numpy().tobytes()) for t in parts))
    h.set_mesh(p.N, p.deg, p.nodes, p.elem_nodes, dof_for_node=p.dof_for_node,
               n_dofs=p.num_dofs if p.dof_for_node is not None else None)
    h.set_interface(p.neighbor_ranks, p.neighbor_offsets, p.shared_local, p.owned)
    h.set_material(D)
    return h


def _volume_form(N, D, vol, strains, cell_volume):
    """sum_e vol_e (rows_i D:strain_i(e) + D) / |Y|   (PeriodicHomogenization.hh:72-100), local part."""
    F = flat_len(N)
    dbl = np.ones(F); dbl[N:] = 2.0
    Eh = np.zeros((F, F))
    for i, s in enumerate(strains):                      # s: (ne, F) average strain of w_i
        Eh[i] += (vol[:, None] * (s * dbl[None, :])).sum(axis=0) @ D.T
    Eh += D * vol.sum()
    return Eh / cell_volume


def homogenize(raw_mesh, deg, D, dist=None, local_rank=0, rtol=1e-10, max_iters=200000, return_fields=False, **options):
    """Homogenized elasticity tensor of the periodic base cell `raw_mesh` (hostlib.RawMesh) with constant
    base material D (flattened).  dist = an initialised torch.distributed module for multi-GPU runs
    (NCCL backend, one process per GPU) or None.  Returns Eh [, dict with the fluctuation displacements
    w_ij on this rank's nodes, the partition and the solver statistics]."""
    info = raw_mesh.apply_bc(deg, "", periodic=True)    # periodic DoFs + the pinned node (m_pinNode)
    m, dfn, nd = info["mesh"], info["dof_for_node"], info["num_dofs"]
    N = m.N
    F = flat_len(N)
    D = np.ascontiguousarray(D, dtype=np.float64)
    cell_volume = float(np.prod(m.bbox_max - m.bbox_min))
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    if world > 1:
        p, lfixed, lvals, _ = local_problem(m, info["fixed_vars"], info["fixed_vals"], None, world, rank, dof_for_node=dfn)
        h = make_handle(dist, world, rank, local_rank, p, D, **options)
        node_dof = p.dof_for_node
    else:
        p = None
        h = Handle(local_rank, **options)
        h.set_mesh(N, deg, m.nodes, m.elem_nodes, dof_for_node=dfn, n_dofs=nd)
        h.set_material(D)
        lfixed, lvals, node_dof = info["fixed_vars"], info["fixed_vals"], dfn
    stats = []
    try:
        h.assemble()
        h.fix_variables(lfixed, lvals)
        w_nodes, strains = [], []
        loads = []
        for i in range(F):
            eps = np.zeros(F); eps[i] = -(1.0 if i < N else 0.5)          # -SMatrix::CanonicalBasis(i)
            loads.append(h.const_strain_load(eps))                          # consistent on shared DoFs
        # flatLen(N) right-hand sides in one call: the library runs them as one batched PCG (SpMM)
        U, stats = h.solve(np.stack(loads), rtol=rtol, max_iters=max_iters, return_info=True)
        for i in range(F):
            w = U[i].reshape(-1, N)[node_dof]                               # dofToNodeField
            w_nodes.append(w)
            strains.append(h.avg_strain_stress(w)[0])
        Eh = _volume_form(N, D, h.volumes(), strains, cell_volume)
    finally:
        h.close()
    if world > 1:
        import torch
        t = torch.from_numpy(Eh.copy())
        if dist.get_backend() == "nccl":
            t = t.cuda(local_rank)
        dist.all_reduce(t)
        Eh = t.cpu().numpy()
        # the constant term D*vol was summed over ranks with the local volumes -> already global
    if return_fields:
        return Eh, dict(w=w_nodes, partition=p, mesh=m, dof_for_node=dfn, solves=stats)
    return Eh
