#!/bin/bash
# Round 2, multi-GPU session 5 (gpurun --gpus 2): overlapped exchange with copy-engine pulls of the FP64 rows.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -x -q -k two_gpu 2>&1 | tail -4
SCEMA_SHARD_OVERLAP=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 --no-library-e2e > gpurun_out/r02_bench_c4_2gpu_pull.json 2> gpurun_out/r02_bench_c4_2gpu_pull.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_c4_2gpu_pull.json")); r=d["roofline"]
print("value %.4g step %.2f ms filter %.2f ms" % (d["value"], d["ms_per_step"], r["launch_ms"]), r["other_kernels_ms"], d["run"]["exchange"])
print("verified", d["verified"]["ok"], d["verified"]["union_equals_single_gpu_list"])
PY
grep -v "^W1\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_bench_c4_2gpu_pull.err | tail -6
