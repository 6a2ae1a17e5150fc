#!/bin/bash
# N-GPU validation: multi-rank parity check (incl. the mid-size golden) + bench at N ranks
N=${1:-4}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1
grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func" gpurun_out/mgpu_check_$N.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 tests/mgpu_two_level_check.py > gpurun_out/mgpu_two_level_$N.log 2>&1
grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func" gpurun_out/mgpu_two_level_$N.log | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $N --steps 5 --warmup 2 > gpurun_out/bench10_${N}gpu.json 2> gpurun_out/bench10_${N}gpu.err
cat gpurun_out/bench10_${N}gpu.json; grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func" gpurun_out/bench10_${N}gpu.err | tail -5
