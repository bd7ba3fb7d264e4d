#!/bin/bash
# Round 2, session 12: header tests after the per-thread registry, bench line sanity.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multirank.py tests/test_dropin_header.py tests/test_cli.py -m gpu -q 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c4_v5.json 2> gpurun_out/r02_bench_c4_v5.err; echo "c4 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_c4_v5.json")); r=d["roofline"]
print("value %.4g step %.2f ms e2e %.2f ms filter %.2f frac %.3f executed %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], r["launch_ms"], r["frac"], r["executed_tflops"]), r["other_kernels_ms"], d["verified"]["ok"], d["roofline_fp64"]["frac"])
PY
