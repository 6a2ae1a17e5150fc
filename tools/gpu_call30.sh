#!/bin/bash
# round-2 final pass: full GPU suite, the driver's bench command, launch list, ncu --set full of the PCG kernels
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/build_smoke30.log 2>&1; tail -1 gpurun_out/build_smoke30.log
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/gpu_tests30.log
cat gpurun_out/gpu_tests30.log
( time timeout 850 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench30_cfg5_1gpu.json 2> gpurun_out/bench30_cfg5_1gpu.err ) 2> gpurun_out/bench30_time.txt
cat gpurun_out/bench30_cfg5_1gpu.json; tail -3 gpurun_out/bench30_cfg5_1gpu.err; cat gpurun_out/bench30_time.txt
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches30_cfg5.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity > gpurun_out/bench_under_ncu30.log 2>&1
tail -2 gpurun_out/bench_under_ncu30.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mf_|k_pcg_update|k_pcg_direction|k_coarse_gemv|k_coarse_level1' -s 12 -c 6 -f \
    -o gpurun_out/pcg30_cfg5 python tools/profile_case.py cfg5 solve coarse_aggregates=2048 > gpurun_out/ncu_pcg30_cfg5.log 2>&1
tail -3 gpurun_out/ncu_pcg30_cfg5.log
ls -la gpurun_out | tail -12
