#!/bin/bash
# Round 2, session 15: ncu --set full of K1 (k_resample_stream) on config 3 (ragged 6..200 steps), all length classes of one step.
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_resample_stream -s 6 -c 2 -o gpurun_out/r02_prof_resample_c3 \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --verify off > gpurun_out/r02_ncu_resample.log 2>&1; echo "ncu rc=$?"
