#!/bin/bash
# Round 2, session 6: bench lines with the verification block: c4, c4s (production-shaped), c4s without centring.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c4_v3.json 2> gpurun_out/r02_bench_c4_v3.err; echo "c4 rc=$?"
timeout 600 python bench.py --workload c4s --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c4s_v1.json 2> gpurun_out/r02_bench_c4s_v1.err; echo "c4s rc=$?"
SCEMA_TC_CENTRE=0 timeout 900 python bench.py --workload c4s --steps 2 --warmup 3 --no-cpu-baseline --verify off > gpurun_out/r02_bench_c4s_nocentre.json 2> gpurun_out/r02_bench_c4s_nocentre.err; echo "c4s nocentre rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_c4_v3","r02_bench_c4s_v1","r02_bench_c4s_nocentre"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d["roofline"]
        print(f, "value %.4g step %.2f ms filter %.2f ms frac %.3f e2e %.2f ms" % (d["value"], d["ms_per_step"], r["launch_ms"], r["frac"] or 0, d["e2e"]["ms_per_step"]), r["other_kernels_ms"])
        print("   run", d["run"]); print("   verified", d["verified"]); print("   fp64", d.get("roofline_fp64")); print("   cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "FAILED", repr(e))
PY
tail -3 gpurun_out/r02_bench_c4s_v1.err
