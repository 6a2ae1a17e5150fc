#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/gpu_tests22.log
cat gpurun_out/gpu_tests22.log
timeout 300 python tools/time_operator.py cfg5 0:8:0:1 0:8:0:3 0:4:0:3 0:8:0:2 > gpurun_out/time_operator22_cfg5.log 2>&1; cat gpurun_out/time_operator22_cfg5.log
( time timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-bj-parity --no-direct > gpurun_out/bench22_cfg5_1gpu.json 2> gpurun_out/bench22_cfg5_1gpu.err ) 2> gpurun_out/bench22_time.txt
cat gpurun_out/bench22_cfg5_1gpu.json; tail -3 gpurun_out/bench22_cfg5_1gpu.err; cat gpurun_out/bench22_time.txt
