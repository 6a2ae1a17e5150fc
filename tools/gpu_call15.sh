#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
for mode in 1 0; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$mode bench.py --gpus $N --steps 5 --warmup 2 --no-bj-parity --comm-p2p $mode > gpurun_out/bench15_${N}gpu_p2p$mode.json 2> gpurun_out/bench15_${N}gpu_p2p$mode.err
done
python - <<PY
import json
for k in ("p2p1","p2p0"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/bench15_${N}gpu_{k}.json") if l.startswith("{")][-1])
        print(k, "value", round(d["value"]), "ms_per_step", round(d["ms_per_step"],2), "solve_ms", round(d["solve_ms"],2), "iters", d["pcg_iterations_per_solve"], "setup", round(d["preconditioner_setup_ms"],2), "asm", round(d["assembly_ms"],2), "e2e_s", round(d["e2e"]["seconds_per_step"],3), "parity", d.get("parity"), d["config"].get("partition","")[-40:])
    except Exception as e:
        print(k, "failed", e)
PY
