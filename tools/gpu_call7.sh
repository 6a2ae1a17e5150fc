#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi -L | head -2
timeout 400 python tools/time_spmv.py cfg5 1:32 1:16 5:0 1:32 5:0 > gpurun_out/spmv7.json 2> gpurun_out/spmv7.err
cat gpurun_out/spmv7.json; tail -3 gpurun_out/spmv7.err
timeout 300 python tools/time_spmv.py cfg3 1:32 1:16 5:0 > gpurun_out/spmv7_cfg3.json 2>> gpurun_out/spmv7.err
cat gpurun_out/spmv7_cfg3.json
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_zz_two_level_gpu.py tests/test_precond_operator_gpu.py -q -m gpu 2>&1 | tail -8 > gpurun_out/gpu_tests7.log
cat gpurun_out/gpu_tests7.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 1 --no-bj-parity > gpurun_out/bench7_2gpu.json 2> gpurun_out/bench7_2gpu.err
cat gpurun_out/bench7_2gpu.json; tail -4 gpurun_out/bench7_2gpu.err
