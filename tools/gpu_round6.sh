#!/bin/bash
# GPU session 6: K1 streamed kernel with in-place ring loads: parity, wps sweep, ncu.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k1 or config3" 2>&1 | tail -6 > gpurun_out/pytest_k1.txt
tail -6 gpurun_out/pytest_k1.txt
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d["roofline"]
    print("%-28s resample_ms=%.4f filter_ms=%.3f TF=%.2f frac=%.3f value=%.4g e2e=%.4g edges=%d" % (sys.argv[2], r["other_kernels_ms"]["resample"], r["launch_ms"], r["achieved"], r["frac"], d["value"], d["e2e"]["value"], d["config"]["edges"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for cfg in "stream 12" "stream 16" "stream 20" "stream 24"; do
  set -- $cfg
  SCEMA_K1=$1 SCEMA_K1_WPS=$2 timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_$1_$2.json 2> gpurun_out/bench_c3_$1_$2.err
  show gpurun_out/bench_c3_$1_$2.json "c3 k1=$1 wps=$2"
done
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
show gpurun_out/bench_c4.json "c4"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resample_stream -s 6 -c 2 -o gpurun_out/prof_resample_v7 \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_resample_v7.log 2>&1
