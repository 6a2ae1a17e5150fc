#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/build_smoke.log 2>&1; tail -1 gpurun_out/build_smoke.log
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/gpu_tests17.log
cat gpurun_out/gpu_tests17.log
( time timeout 850 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench17_cfg5_1gpu.json 2> gpurun_out/bench17_cfg5_1gpu.err ) 2> gpurun_out/bench17_time.txt
cat gpurun_out/bench17_cfg5_1gpu.json; tail -3 gpurun_out/bench17_cfg5_1gpu.err; cat gpurun_out/bench17_time.txt
timeout 300 python bench.py --config cfg3 --steps 5 --warmup 2 --no-direct > gpurun_out/bench17_cfg3.json 2> gpurun_out/bench17_cfg3.err
cat gpurun_out/bench17_cfg3.json
timeout 300 python bench.py --config cfg2 --steps 5 --warmup 2 --no-direct --coarse-aggregates=-1 > gpurun_out/bench17_cfg2.json 2> gpurun_out/bench17_cfg2.err
cat gpurun_out/bench17_cfg2.json; tail -2 gpurun_out/bench17_cfg2.err
COARSE=-1 timeout 400 python tools/homog_bench.py 64 2 > gpurun_out/homog17_cfg4_multilevel.json 2> gpurun_out/homog17.err
cat gpurun_out/homog17_cfg4_multilevel.json; tail -2 gpurun_out/homog17.err
( time timeout 850 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench17_reference.json 2> gpurun_out/bench17_reference.err ) 2> gpurun_out/bench17_ref_time.txt
cat gpurun_out/bench17_reference.json; cat gpurun_out/bench17_ref_time.txt
