#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $N --steps 5 --warmup 2 > gpurun_out/bench11_${N}gpu.json 2> gpurun_out/bench11_${N}gpu.err
cat gpurun_out/bench11_${N}gpu.json; grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func" gpurun_out/bench11_${N}gpu.err | tail -5
