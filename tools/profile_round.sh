#!/bin/bash
# Round profile pass on the GPU box (run through gpurun): default bench, ncu launch list of the
# same command, and one `ncu --set full` capture of the PCG kernels and of the assembly kernel.
# Everything lands in gpurun_out/; summaries are copied into profiles/ by tools/summarise_ncu.py.
set -u
mkdir -p gpurun_out
CFG=${1:-cfg5}
timeout 900 python bench.py --config $CFG > gpurun_out/bench_$CFG.json 2> gpurun_out/bench_$CFG.err
tail -c 3000 gpurun_out/bench_$CFG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$CFG.csv \
    python bench.py --config $CFG --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_under_ncu_$CFG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_bsr_spmv|k_pcg_update|k_pcg_direction' -s 30 -c 6 -f \
    -o gpurun_out/pcg_$CFG python tools/profile_case.py $CFG solve > gpurun_out/ncu_pcg_$CFG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_blocks -c 1 -f \
    -o gpurun_out/asm_$CFG python tools/profile_case.py $CFG assemble > gpurun_out/ncu_asm_$CFG.log 2>&1
ls -la gpurun_out
