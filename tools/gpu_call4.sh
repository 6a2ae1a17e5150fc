#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_precond_operator_gpu.py tests/test_zz_two_level_gpu.py -q -m gpu 2>&1 | tail -30 > gpurun_out/gpu_tests4.log
cat gpurun_out/gpu_tests4.log
timeout 300 python tools/precond_sweep.py --max-iters 1500 --config cfg3 --combos 2048:0,2048:32,2048:16,1024:32,-1:32 > gpurun_out/sweep_cfg3.json 2> gpurun_out/sweep_cfg3.err
cat gpurun_out/sweep_cfg3.json; tail -3 gpurun_out/sweep_cfg3.err
timeout 600 python tools/precond_sweep.py --max-iters 1500 --config cfg5 --combos 2048:0,2048:32,2048:64,2048:16,4096:32,1024:32 > gpurun_out/sweep_cfg5.json 2> gpurun_out/sweep_cfg5.err
cat gpurun_out/sweep_cfg5.json; tail -3 gpurun_out/sweep_cfg5.err
timeout 200 python tools/precond_sweep.py --max-iters 1500 --config cfg2 --combos 0:0,-1:0,-1:32,-1:16 > gpurun_out/sweep_cfg2.json 2> gpurun_out/sweep_cfg2.err
cat gpurun_out/sweep_cfg2.json; tail -3 gpurun_out/sweep_cfg2.err
