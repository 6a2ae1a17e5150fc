"""CPU: bench.py's control flow and JSON line with a mock Handle (tests/bench_dryrun_worker.py), in a subprocess because
the worker replaces meshfem_b200.Handle."""
import os
import subprocess
import sys

from util import ROOT


def test_bench_json_line_with_mock_handle(lib_built):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_dryrun_worker.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "DRYRUN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l[:2] in ("-1", "0 ")]
    assert "k_mf_chunk" in lines[0] and "fp64" in lines[0] and "k_bsr_spmv" in lines[1]
