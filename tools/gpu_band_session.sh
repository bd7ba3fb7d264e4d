#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x --durations=5 2>&1 | tail -n 12
SCEMA_TC_SLICES=1 timeout 300 python tools/tc_probe.py 1000 1000000 2>&1 | grep "big n" | sed -n 2p
timeout 600 python bench.py --no-cpu-baseline --norm-band > gpurun_out/bench_c4_band.json 2> gpurun_out/bench_c4_band.err; tail -2 gpurun_out/bench_c4_band.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c4_dense.json 2> /dev/null
python - <<'PY'
import json
for f in ("bench_c4_band","bench_c4_dense"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d["roofline"]
        print(f, "value %.4g ms %.3f e2e %.4g ms %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), "filter_ms %.3f" % r["launch_ms"], "band_tiles", d["config"].get("band_tiles_last_rank0"), "edges", d["config"]["edges"], r["other_kernels_ms"])
    except Exception as e: print(f, "FAILED", e)
PY
