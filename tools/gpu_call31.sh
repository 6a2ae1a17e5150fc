#!/bin/bash
mkdir -p gpurun_out
N=2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/mgpu_check.py > gpurun_out/mgpu_check31_$N.log 2>&1
tail -3 gpurun_out/mgpu_check31_$N.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $N --steps 3 --warmup 2 --no-bj-parity > gpurun_out/bench31_${N}gpu.json 2> gpurun_out/bench31_${N}gpu.err
cat gpurun_out/bench31_${N}gpu.json; tail -2 gpurun_out/bench31_${N}gpu.err
