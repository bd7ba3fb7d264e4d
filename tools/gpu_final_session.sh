#!/bin/bash
# Final single-GPU session of the round: whole GPU suite, smoke, both bench arms, c3 / c2 / c5-shape lines, launch
# list of the default bench command, ncu --set full captures of the dominant kernel, compute-sanitizer on the small
# tcgen05 cases.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest_gpu_final.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu_final.log; tail -n 10 gpurun_out/pytest_gpu_final.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> /dev/null
timeout 600 python bench.py > gpurun_out/bench_c4_final.json 2> gpurun_out/bench_c4_final.err; cat gpurun_out/bench_c4_final.json
timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_final.json 2> /dev/null
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_final.json 2> /dev/null
timeout 600 python bench.py --workload c5 --histories 300000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_300k_final.json 2> /dev/null
SCEMA_TC_CG=2 timeout 600 python bench.py --workload c5 --histories 300000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_300k_cg2.json 2> /dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c4_final.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_filter_tc -s 3 -c 1 -o gpurun_out/prof_filter_tc_final \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_filter_tc_final.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "accumulators and (clusters or k18) or wide_rows and 50 or norm_band" > gpurun_out/sanitizer_memcheck_tc.txt 2>&1; echo "memcheck rc=$?"
python - <<'PY'
import json
for f in ("bench_ref_final","bench_c4_final","bench_c3_final","bench_c2_final","bench_c5_300k_final","bench_c5_300k_cg2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d.get("roofline") or {}
        print(f, "value %.4g ms %.2f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "filter_ms", r.get("launch_ms"), "frac", r.get("frac"), r.get("other_kernels_ms"))
    except Exception as e: print(f, "FAILED", e)
PY
tail -5 gpurun_out/sanitizer_memcheck_tc.txt
