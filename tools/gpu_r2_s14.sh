#!/bin/bash
# Round 2, session 14: TMEM probe (fixed two-loads mode), whole GPU suite, smoke, both bench arms as the driver runs them.
mkdir -p gpurun_out
timeout 200 ./tools/tmem_probe > gpurun_out/r02_tmem_probe_v2.json 2> gpurun_out/r02_tmem_probe_v2.err; echo "probe rc=$?"
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -n 10 gpurun_out/r02_pytest_final.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_final.json 2>/dev/null; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_bench_c4_final.json 2> gpurun_out/r02_bench_c4_final.err; echo "c4 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_tmem_probe_v2.json"))
for r in d["results"]:
    if r["warps"]==8: print(r)
d=json.load(open("gpurun_out/r02_bench_c4_final.json")); r=d["roofline"]
print("value %.4g step %.2f ms e2e %.2f ms filter %.2f frac %.3f executed %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], r["launch_ms"], r["frac"], r["executed_tflops"]), r["other_kernels_ms"], d["verified"]["ok"], d["roofline_fp64"]["frac"], d["cpu_baseline"]["value"])
PY
