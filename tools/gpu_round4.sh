#!/bin/bash
# GPU session 4: K1 streamed kernel parity + A/B timing against the staged kernel, WPS sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k1 or config3 or smoke" 2>&1 | tail -15 > gpurun_out/pytest_k1.txt
tail -15 gpurun_out/pytest_k1.txt
for cfg in "stream 8" "stream 12" "stream 16" "stream 24" "staged 16"; do
  set -- $cfg
  SCEMA_K1=$1 SCEMA_K1_WPS=$2 timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_$1_$2.json 2> gpurun_out/bench_c3_$1_$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_c3_$1_$2.json"))
print("$1 wps=$2 resample_ms=%.4f filter_ms=%.3f value=%.4g e2e=%.4g" % (d["roofline"]["other_kernels_ms"]["resample"], d["roofline"]["launch_ms"], d["value"], d["e2e"]["value"]))
PY
done
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt
tail -8 gpurun_out/pytest_gpu.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resample_stream -s 6 -c 2 -o gpurun_out/prof_resample_v4 \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_resample_v4.log 2>&1
ls gpurun_out
