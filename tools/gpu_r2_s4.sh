#!/bin/bash
# Round 2, session 4: A/B of CG=1/2 after the epilogue AND change, config 4 and the config-5 shape.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/r02_pytest_s4.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02_pytest_s4.log
for cfg in "c4 1 1000000" "c4 2 1000000" "c5 1 300000" "c5 2 300000"; do
  set -- $cfg
  SCEMA_TC_CG=$2 timeout 300 python bench.py --workload $1 --histories $3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_$1_cg$2_v2.json 2> gpurun_out/r02_bench_$1_cg$2_v2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_$1_cg$2_v2.json")); r=d["roofline"]
print("$cfg", "step %.2f ms filter %.2f ms frac %.3f edges %d e2e %.2f ms" % (d["ms_per_step"], r["launch_ms"], r["frac"], d["config"]["edges"], d["e2e"]["ms_per_step"]), r["other_kernels_ms"])
PY
done
