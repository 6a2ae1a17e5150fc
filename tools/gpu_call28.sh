#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_matrix_free_gpu.py -q -x 2>&1 | tail -25 > gpurun_out/mf_tests28.log
cat gpurun_out/mf_tests28.log
timeout 400 python tools/time_operator.py cfg5 1:0:4:0:3:64:16 1:0:4:0:3:128:16 1:0:4:0:3:32:16 1:0:4:0:3:64:12 > gpurun_out/time_operator28_cfg5.log 2>&1; cat gpurun_out/time_operator28_cfg5.log
