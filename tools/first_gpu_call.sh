#!/bin/bash
# First GPU call of the next round (everything below was written after round 1's GPU budget was spent and has
# never run on a device).  One box, ~12 minutes:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
# Results land in gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
# 1. the never-run GPU tests, all of them (no -x), slowest-to-diagnose first
timeout 1200 python -m pytest tests/test_zz_two_level_gpu.py tests/test_zz_lagrange_gpu.py tests/test_zz_deformed_cells_gpu.py \
    tests/test_zz_periodic_variants_gpu.py tests/test_zz_shape_derivatives_gpu.py tests/test_zy_microstructures_gpu.py \
    -q -m gpu 2>&1 | tail -80 > gpurun_out/zz_tests.log
# 2. the two-level preconditioner on the bench workloads (iterations, step time, e2e, validation)
for cfg_s in "cfg3 1024 0" "cfg3 1024 1" "cfg5 2048 0" "cfg5 4096 0" "cfg5 2048 1"; do
    set -- $cfg_s
    timeout 300 python tools/two_level_trial.py --config $1 --aggregates $2 --shape $3 > gpurun_out/two_level_$1_$2_shape$3.json 2> gpurun_out/two_level_$1_$2_shape$3.err
done
# 3. where the coarse-space time goes: launch list of one trial (per-launch times are cold-cache and serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_two_level_cfg3.csv \
    python tools/two_level_trial.py --config cfg3 --aggregates 1024 > gpurun_out/ncu_trial.log 2>&1
# 4. host side: threaded FEMMesh build on real host cores (opt-in until measured)
for t in 1 4 8 16; do
    MESHFEM_NUM_THREADS=$t python - <<'PY' >> gpurun_out/femmesh_threads.log 2>&1
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from meshfem_b200 import hostlib
raw = hostlib.grid([220, 44, 44])
t = time.perf_counter(); m = raw.femmesh(2)
print(os.environ["MESHFEM_NUM_THREADS"], "threads: femmesh cfg5", round(time.perf_counter() - t, 2), "s", m.num_elements)
PY
done
tail -5 gpurun_out/zz_tests.log
cat gpurun_out/two_level_*.json
cat gpurun_out/femmesh_threads.log
