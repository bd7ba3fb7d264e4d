#!/bin/bash
# Round 2, session 5: centred filter copies + up-front filter choice — parity, then the bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tc.py -m gpu -q --durations=5 > gpurun_out/r02_pytest_s5.log 2>&1; echo "pytest rc=$?"; tail -n 40 gpurun_out/r02_pytest_s5.log | cut -c1-400
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c4_s5.json 2> gpurun_out/r02_bench_c4_s5.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_c4_s5.json")); r=d["roofline"]
print("c4", "step %.2f ms filter %.2f ms frac %.3f edges %d e2e %.2f ms" % (d["ms_per_step"], r["launch_ms"], r["frac"], d["config"]["edges"], d["e2e"]["ms_per_step"]), r["other_kernels_ms"], d["config"]["survivors_last_rank0"])
PY
