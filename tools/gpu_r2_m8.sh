#!/bin/bash
# Round 2, 8-GPU session (gpurun --gpus 8): config 4 with verification + library e2e, config 5 at full size with the four
# checks of SURVEY 8d (--verify full), multi-GPU tests.
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_c4_8gpu_v1.json 2> gpurun_out/r02_bench_c4_8gpu_v1.err
echo "c4 rc=$?"
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --workload c5 --steps 2 --warmup 3 --verify full > gpurun_out/r02_bench_c5_full_8gpu_v1.json 2> gpurun_out/r02_bench_c5_full_8gpu_v1.err
echo "c5 rc=$?"
timeout 900 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -3
python - <<'PY'
import json
for f in ("r02_bench_c4_8gpu_v1", "r02_bench_c5_full_8gpu_v1"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d["roofline"]
        print(f, "value %.4g step %.2f ms filter %.2f ms" % (d["value"], d["ms_per_step"], r["launch_ms"]), r["other_kernels_ms"])
        print("  e2e", d["e2e"]); print("  e2e_ppg", d.get("e2e_process_per_gpu")); print("  verified", d["verified"]); print("  run", d["run"])
    except Exception as e: print(f, "FAILED", repr(e))
PY
grep -v "^W1\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_bench_c5_full_8gpu_v1.err | tail -5
