#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k1 or config3 or store" 2>&1 | tail -5
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print("%-22s value=%.4g ms/step=%.2f e2e=%.4g filter_ms=%.3f TF=%.2f other=%s" % (sys.argv[2], d["value"], d["ms_per_step"], d["e2e"]["value"], r["launch_ms"], r["achieved"], {k: round(v,4) for k,v in r["other_kernels_ms"].items()}))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for wps in 10 12 16 20 24; do
SCEMA_K1_WPS=$wps timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_w$wps.json 2> gpurun_out/bench_c3_w$wps.err; show gpurun_out/bench_c3_w$wps.json "c3 wps=$wps"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resample_stream -s 3 -c 2 -o gpurun_out/prof_resample_v10 \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_resample_v10.log 2>&1
