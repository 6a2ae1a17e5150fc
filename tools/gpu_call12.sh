#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/build_smoke.log 2>&1; tail -1 gpurun_out/build_smoke.log
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/gpu_tests12.log
cat gpurun_out/gpu_tests12.log
timeout 300 python tools/precond_sweep.py --max-iters 3000 --config cfg2 --combos 0:0,-1:64 > gpurun_out/sweep12_cfg2.json 2> gpurun_out/sweep12.err
cat gpurun_out/sweep12_cfg2.json
timeout 300 python tools/precond_sweep.py --max-iters 3000 --config cfg3 --combos -1:64 > gpurun_out/sweep12_cfg3.json 2>> gpurun_out/sweep12.err
cat gpurun_out/sweep12_cfg3.json; tail -3 gpurun_out/sweep12.err
