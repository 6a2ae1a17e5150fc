#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_zz_two_level_gpu.py -q -m gpu -k "ranks or multi" 2>&1 | tail -30 > gpurun_out/gpu_tests6.log
cat gpurun_out/gpu_tests6.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench6_2gpu.json 2> gpurun_out/bench6_2gpu.err
cat gpurun_out/bench6_2gpu.json; tail -8 gpurun_out/bench6_2gpu.err
