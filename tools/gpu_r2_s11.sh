#!/bin/bash
# Round 2, session 11: whole GPU suite, smoke, bench lines (c4 default, c4s), launch list of the default bench command.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r02_pytest_s11.log 2>&1; echo "pytest rc=$?"; tail -n 14 gpurun_out/r02_pytest_s11.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-300
timeout 600 python bench.py > gpurun_out/r02_bench_c4_v4.json 2> gpurun_out/r02_bench_c4_v4.err; echo "c4 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_v1.json 2> /dev/null; echo "ref rc=$?"
timeout 600 python bench.py --workload c4s --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c4s_v2.json 2> gpurun_out/r02_bench_c4s_v2.err; echo "c4s rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_c4.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --verify off > gpurun_out/r02_ncu_launch.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_c4_v4","r02_bench_c4s_v2","r02_bench_ref_v1"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d.get("roofline") or {}
        print(f, "value %.4g step %.2f ms e2e %.4g (%.2f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step", 0)), "filter_ms", r.get("launch_ms"), "frac", r.get("frac"), r.get("other_kernels_ms"), "verified", (d.get("verified") or {}).get("ok"), "fp64", (d.get("roofline_fp64") or {}).get("frac"))
    except Exception as e: print(f, "FAILED", repr(e))
PY
