#!/bin/bash
# Round 2, multi-GPU session 1 (gpurun --gpus 2): sharded parity incl. the tcgen05 filter, bench with verification.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_c4_2gpu_v1.json 2> gpurun_out/r02_bench_c4_2gpu_v1.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_c4_2gpu_v1.json")); r=d["roofline"]
print("value %.4g step %.2f ms filter %.2f ms e2e %.2f ms" % (d["value"], d["ms_per_step"], r["launch_ms"], d["e2e"]["ms_per_step"]), r["other_kernels_ms"])
print("verified", d["verified"])
PY
grep -v "^W1\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_bench_c4_2gpu_v1.err | tail -5
