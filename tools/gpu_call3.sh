#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 300 python tools/debug_coarse.py 2>&1 | tail -80 > gpurun_out/debug_coarse.log
cat gpurun_out/debug_coarse.log
