#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
tail -8 gpurun_out/pytest_gpu.txt
timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench.err; tail -c 1200 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resample_staged -s 12 -c 3 -o gpurun_out/prof_resample \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_resample.log 2>&1
