#!/bin/bash
# GPU session 3: parity tests on the restored tree, FP64 mixed-issue probe, K=300 (config-5 shape) bench + ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
tail -8 gpurun_out/pytest_gpu.txt
./tools/fp64_peak > gpurun_out/fp64_peak.json; cat gpurun_out/fp64_peak.json
timeout 600 python bench.py --workload c5 --histories 300000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5s.json 2> gpurun_out/bench_c5s.err; tail -c 2500 gpurun_out/bench_c5s.json; tail -3 gpurun_out/bench_c5s.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_filter -s 3 -c 1 -o gpurun_out/prof_filter_k300 \
    python bench.py --workload c5 --histories 100000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_filter_k300.log 2>&1
ls -la gpurun_out
