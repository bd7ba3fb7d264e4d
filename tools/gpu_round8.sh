#!/bin/bash
# GPU session 8: history store, streaming compare, c5-shape streaming bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --workload c5 --histories 300000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5s_stream.json 2> gpurun_out/bench_c5s_stream.err; tail -c 1800 gpurun_out/bench_c5s_stream.json; tail -3 gpurun_out/bench_c5s_stream.err
timeout 600 python bench.py --workload c5 --histories 300000 --steps 3 --warmup 3 --no-cpu-baseline --stream 0 > gpurun_out/bench_c5s_oneshot.json 2> gpurun_out/bench_c5s_oneshot.err; tail -c 600 gpurun_out/bench_c5s_oneshot.json
