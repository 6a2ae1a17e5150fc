#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/build_smoke.log 2>&1
timeout 600 python -m pytest tests/test_zz_two_level_gpu.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -40 > gpurun_out/gpu_tests2.log
tail -15 gpurun_out/gpu_tests2.log
timeout 400 python tools/precond_sweep.py --config cfg3 --combos 2048:0,2048:32,2048:16,-1:32 > gpurun_out/sweep_cfg3.json 2> gpurun_out/sweep_cfg3.err
cat gpurun_out/sweep_cfg3.json; tail -3 gpurun_out/sweep_cfg3.err
timeout 600 python tools/precond_sweep.py --config cfg5 --combos 2048:0,2048:32,2048:64,2048:16,4096:32,1024:32 > gpurun_out/sweep_cfg5.json 2> gpurun_out/sweep_cfg5.err
cat gpurun_out/sweep_cfg5.json; tail -3 gpurun_out/sweep_cfg5.err
timeout 200 python tools/precond_sweep.py --config cfg2 --combos 0:0,-1:0,-1:32,-1:16 > gpurun_out/sweep_cfg2.json 2> gpurun_out/sweep_cfg2.err
cat gpurun_out/sweep_cfg2.json; tail -3 gpurun_out/sweep_cfg2.err
