#!/bin/bash
# One GPU session: parity tests, bench (both arms), ncu launch list + full capture of the K2 filter.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 3000 gpurun_out/bench_c4.json; tail -5 gpurun_out/bench_c4.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2>> gpurun_out/bench_c4.err; tail -c 2500 gpurun_out/bench_c3.json
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2>> gpurun_out/bench_c4.err; tail -c 1500 gpurun_out/bench_c2.json
# launch list (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
# full capture of the dominant kernel and of K1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_filter -s 3 -c 1 -o gpurun_out/prof_filter \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_filter.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resample_staged -s 12 -c 2 -o gpurun_out/prof_resample \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_resample.log 2>&1
ls -la gpurun_out
