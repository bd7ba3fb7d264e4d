#!/bin/bash
# Round 2, session 2: tcgen05 filter with warp-uniform issue loops — parity, then A/B against the round-1 library.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -x -q --durations=5 > gpurun_out/r02_pytest_s2.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r02_pytest_s2.log
for cfg in "new 1 1" "new 2 1" "new 1 2" "r1 1 1"; do
  set -- $cfg
  L=$PWD/scema_b200/libscema_hist.so; [ $1 != new ] && L=$PWD/scema_b200/libscema_hist_$1.so
  SCEMA_LIB=$L SCEMA_TC_CG=$2 SCEMA_TC_SLICES=$3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c4_$1_cg$2_s$3.json 2> gpurun_out/r02_bench_c4_$1_cg$2_s$3.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_c4_$1_cg$2_s$3.json")); r=d["roofline"]
print("$cfg", "step %.2f ms filter %.2f ms frac %.3f edges %d e2e %.2f ms" % (d["ms_per_step"], r["launch_ms"], r["frac"], d["config"]["edges"], d["e2e"]["ms_per_step"]), r["other_kernels_ms"])
PY
done
