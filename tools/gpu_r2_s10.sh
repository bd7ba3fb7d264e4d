#!/bin/bash
# Round 2, session 10: config 5 at FULL size (4M histories x 300 columns, 8.0e12 pairs, edges streamed) on ONE GPU with the
# four checks of SURVEY 8d (bench.py --verify full).
mkdir -p gpurun_out
timeout 840 python bench.py --workload c5 --steps 2 --warmup 3 --verify full --no-cpu-baseline > gpurun_out/r02_bench_c5_full_1gpu_v1.json 2> gpurun_out/r02_bench_c5_full_1gpu_v1.err
echo "c5 rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02_bench_c5_full_1gpu_v1.json")); r=d["roofline"]
    print("value %.4g step %.2f ms filter %.2f ms frac %.3f" % (d["value"], d["ms_per_step"], r["launch_ms"], r["frac"]), r["other_kernels_ms"])
    print("  e2e", d["e2e"]); print("  verified", d["verified"]); print("  run", d["run"])
except Exception as e: print("FAILED", repr(e))
PY
tail -3 gpurun_out/r02_bench_c5_full_1gpu_v1.err
