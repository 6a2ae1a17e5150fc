#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 400 python tools/time_spmv.py cfg5 1:32:0 1:32:1 1:16:1 1:32:0 1:32:1 > gpurun_out/spmv8.json 2> gpurun_out/spmv8.err
cat gpurun_out/spmv8.json; tail -3 gpurun_out/spmv8.err
timeout 300 python tools/time_spmv.py cfg3 1:32:0 1:32:1 1:16:1 > gpurun_out/spmv8_cfg3.json 2>> gpurun_out/spmv8.err
cat gpurun_out/spmv8_cfg3.json
timeout 300 python tools/time_spmv.py cfg2 1:8:0 1:8:1 1:16:0 1:16:1 > gpurun_out/spmv8_cfg2.json 2>> gpurun_out/spmv8.err
cat gpurun_out/spmv8_cfg2.json
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "spmv" 2>&1 | tail -5 > gpurun_out/gpu_tests8.log
cat gpurun_out/gpu_tests8.log
