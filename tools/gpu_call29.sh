#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_matrix_free_gpu.py -q -x 2>&1 | tail -25 > gpurun_out/mf_tests29.log
cat gpurun_out/mf_tests29.log
timeout 400 python tools/time_operator.py cfg5 1:0:4:0:3:64:16 1:0:1:0:3:64:16 1:0:1:0:3:64:20 1:0:1:0:3:128:20 > gpurun_out/time_operator29_cfg5.log 2>&1; cat gpurun_out/time_operator29_cfg5.log
