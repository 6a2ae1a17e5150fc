#!/bin/bash
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:'k_mf_' -s 2 -c 2 -f \
    -o gpurun_out/mf2_cfg5 python tools/profile_case.py cfg5 solve matrix_free=1 coarse_aggregates=2048 > gpurun_out/ncu_mf2_cfg5.log 2>&1
tail -3 gpurun_out/ncu_mf2_cfg5.log
