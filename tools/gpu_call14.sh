#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/mgpu_check.py > gpurun_out/mgpu_check14_$N.log 2>&1
grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func" gpurun_out/mgpu_check14_$N.log | grep "grid\|periodic\|MGPU\|Error" | tail -14
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $N --steps 3 --warmup 1 --no-bj-parity > gpurun_out/bench14_${N}gpu_p2p.json 2> gpurun_out/bench14_${N}gpu_p2p.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29554 bench.py --gpus $N --steps 3 --warmup 1 --no-bj-parity --no-parity --comm-p2p 0 > gpurun_out/bench14_${N}gpu_nccl.json 2> gpurun_out/bench14_${N}gpu_nccl.err
python - <<PY
import json
for k in ("p2p","nccl"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/bench14_${N}gpu_{k}.json") if l.startswith("{")][-1])
        print(k, "ms_per_step", round(d["ms_per_step"],2), "solve_ms", round(d["solve_ms"],2), "iters", d["pcg_iterations_per_solve"], "setup", round(d["preconditioner_setup_ms"],2), "e2e_s", round(d["e2e"]["seconds_per_step"],3), "parity", d.get("parity"), d["config"].get("partition","")[-40:])
    except Exception as e:
        print(k, "failed", e)
PY
grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func" gpurun_out/bench14_${N}gpu_p2p.err | grep -i "error" | head -3
