#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/build_smoke.log 2>&1; tail -2 gpurun_out/build_smoke.log
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/gpu_tests5.log
cat gpurun_out/gpu_tests5.log
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench5.json 2> gpurun_out/bench5.err
cat gpurun_out/bench5.json; tail -5 gpurun_out/bench5.err
timeout 600 python tools/precond_sweep.py --max-iters 1500 --config cfg5 --combos 2048:64,2048:128,2048:256,3072:64,1536:96 > gpurun_out/sweep5_cfg5.json 2> gpurun_out/sweep5_cfg5.err
cat gpurun_out/sweep5_cfg5.json; tail -3 gpurun_out/sweep5_cfg5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches5_cfg5.csv \
    python tools/profile_case.py cfg5 solve coarse_aggregates=2048 coarse_fine_nodes=64 > gpurun_out/ncu_launch5.log 2>&1
tail -3 gpurun_out/ncu_launch5.log
