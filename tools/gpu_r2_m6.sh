#!/bin/bash
# Round 2, multi-GPU session 6 (gpurun --gpus 4): the default bench command at N = 4 (never run this round), both arms.
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29573 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_4gpu.json 2>/dev/null; echo "ref rc=$?"; head -c 300 gpurun_out/r02_bench_ref_4gpu.json; echo
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29574 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_c4_4gpu_v1.json 2> gpurun_out/r02_bench_c4_4gpu_v1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_c4_4gpu_v1.json")); r=d["roofline"]
print("value %.4g step %.2f ms filter %.2f ms" % (d["value"], d["ms_per_step"], r["launch_ms"]), r["other_kernels_ms"])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["phases_ms"]); print("verified", d["verified"]["ok"], d["verified"]["sum_keys_mod_2_64"], d["verified"]["xor_distance_bits"])
PY
grep -v "^W1\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_bench_c4_4gpu_v1.err | tail -4
