#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 300 python tools/time_spmv.py cfg5 1:32:0:0 1:32:0:1 1:32:0:0 1:32:0:1 > gpurun_out/spmv16.json 2> gpurun_out/spmv16.err
cat gpurun_out/spmv16.json; tail -2 gpurun_out/spmv16.err
timeout 200 python tools/time_spmv.py cfg3 1:32:0:0 1:32:0:1 > gpurun_out/spmv16_cfg3.json 2>> gpurun_out/spmv16.err
cat gpurun_out/spmv16_cfg3.json
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_shape_kernels_gpu.py -q -m gpu 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_bsr_spmv|k_pcg_update|k_pcg_direction|k_coarse_gemv|k_coarse_level1' -s 10 -c 10 -f \
    -o gpurun_out/pcg16_cfg5 python tools/profile_case.py cfg5 solve coarse_aggregates=2048 coarse_fine_nodes=64 > gpurun_out/ncu_pcg16.log 2>&1
tail -2 gpurun_out/ncu_pcg16.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_blocks|k_coarse_matrix|k_coarse_diag1' -c 3 -f \
    -o gpurun_out/asm16_cfg5 python tools/profile_case.py cfg5 solve coarse_aggregates=2048 coarse_fine_nodes=64 > gpurun_out/ncu_asm16.log 2>&1
tail -2 gpurun_out/ncu_asm16.log
ls -la gpurun_out/*.ncu-rep
