#!/bin/bash
# 8-GPU session (gpurun --gpus 8): sharded parity check, config 5 at full size (4M x 300) streamed, config 4.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --workload c5 --steps 1 --warmup 3 > gpurun_out/bench_c5_full_8gpu_tc.json 2> gpurun_out/bench_c5_full_8gpu_tc.err
echo "exit $?"; cat gpurun_out/bench_c5_full_8gpu_tc.json; grep -v "^W1\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_c5_full_8gpu_tc.err | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_c4_8gpu_tc.json 2> gpurun_out/bench_c4_8gpu_tc.err
echo "exit $?"; cat gpurun_out/bench_c4_8gpu_tc.json
