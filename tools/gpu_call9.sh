#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 400 python tools/time_spmv.py cfg5 1:32:0:0 1:32:0:3 1:32:0:5 1:32:1:0 1:32:0:0 1:32:1:0 > gpurun_out/spmv9.json 2> gpurun_out/spmv9.err
cat gpurun_out/spmv9.json; tail -3 gpurun_out/spmv9.err
timeout 300 python tools/time_spmv.py cfg3 1:32:0:0 1:32:0:3 1:32:0:5 1:32:1:0 > gpurun_out/spmv9_cfg3.json 2>> gpurun_out/spmv9.err
cat gpurun_out/spmv9_cfg3.json
timeout 900 python -m pytest tests/test_parity_midsize_gpu.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -12 > gpurun_out/gpu_tests9.log
cat gpurun_out/gpu_tests9.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-bj-parity --no-direct > gpurun_out/bench9.json 2> gpurun_out/bench9.err
cat gpurun_out/bench9.json; tail -3 gpurun_out/bench9.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_bsr_spmv|k_pcg|k_coarse_level1|k_coarse_gemv' -c 300 --csv --log-file gpurun_out/launches9_iter_cfg5.csv \
    python tools/profile_case.py cfg5 solve coarse_aggregates=2048 coarse_fine_nodes=64 > gpurun_out/ncu_launch9.log 2>&1
tail -2 gpurun_out/ncu_launch9.log
