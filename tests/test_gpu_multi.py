"""N>1 on real GPUs (skipped when fewer than 2 devices are visible): tests/mgpu_check.py under
torchrun -- every rank's slab solution against the oracle's direct solve, 1e-8 relative L2."""
import os
import subprocess
import sys

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_solve_matches_direct_solve(lib_built):
    import meshfem_b200
    n = meshfem_b200.load_library().mfem_b200_device_count()
    if n < 2:
        pytest.skip(f"{n} CUDA device(s) visible; the world_size-2 gloo test covers the host logic on CPU")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MGPU_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
