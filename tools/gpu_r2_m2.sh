#!/bin/bash
# Round 2, multi-GPU session 2 (gpurun --gpus 2): C-ABI multi-GPU path, CLI on two GPUs, bench with library e2e.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_c4_2gpu_v4.json 2> gpurun_out/r02_bench_c4_2gpu_v4.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_c4_2gpu_v4.json")); r=d["roofline"]
print("value %.4g step %.2f ms filter %.2f ms" % (d["value"], d["ms_per_step"], r["launch_ms"]), r["other_kernels_ms"])
print("e2e", d["e2e"]); print("e2e_process_per_gpu", d["e2e_process_per_gpu"])
print("verified", d["verified"])
PY
grep -v "^W1\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_bench_c4_2gpu_v4.err | tail -8
