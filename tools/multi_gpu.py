"""Multi-GPU driver pieces shared by bench.py and the N>1 tests: one process per GPU
(torchrun), torch.distributed for rendezvous / barriers / max-over-ranks, the library's own NCCL
communicator (created from a unique id broadcast through torch.distributed) for the data path."""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


from meshfem_b200.distributed import broadcast_bytes, local_problem  # noqa: E402,F401
from meshfem_b200 import distributed as _dist_mod  # noqa: E402


def make_handle(mfem, dist, world, rank, local_rank, p, D, comm_parent=None, **options):
    """Handle with communicator, local mesh, interface and material (meshfem_b200.distributed)."""
    return _dist_mod.make_handle(dist, world, rank, local_rank, p, D, comm_parent=comm_parent, **options)


def max_over_ranks(dist, value, device):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, value, device):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
