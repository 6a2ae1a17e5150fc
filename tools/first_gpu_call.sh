#!/bin/bash
# First GPU call of a round: the whole GPU suite WITHOUT -x (one stale assertion must not hide the rest), smoke, and the
# two-level preconditioner on the bench workloads.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/first_gpu_call.sh'
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/build_smoke.log 2>&1
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -80 > gpurun_out/gpu_tests.log
for cfg_s in "cfg5 2048 0" "cfg5 4096 0" "cfg5 5400 0" "cfg3 2048 0" "cfg2 2048 0"; do
    set -- $cfg_s
    timeout 200 python tools/two_level_trial.py --config $1 --aggregates $2 --shape $3 > gpurun_out/two_level_$1_$2_shape$3.json 2> gpurun_out/two_level_$1_$2_shape$3.err
done
tail -5 gpurun_out/build_smoke.log
tail -30 gpurun_out/gpu_tests.log
cat gpurun_out/two_level_*.json
