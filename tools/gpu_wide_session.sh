#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x --durations=5 > gpurun_out/pytest_tc.log 2>&1
echo "exit $?" >> gpurun_out/pytest_tc.log; tail -n 25 gpurun_out/pytest_tc.log
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 --deselect tests/test_gpu_tc.py > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log; tail -n 12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload c5 --histories 300000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_300k_tc.json 2> gpurun_out/bench_c5_300k_tc.err
echo "exit $?" >> gpurun_out/bench_c5_300k_tc.err; tail -3 gpurun_out/bench_c5_300k_tc.err
timeout 600 python bench.py --workload c5 --histories 300000 --steps 3 --warmup 3 --no-cpu-baseline --stream 0 > gpurun_out/bench_c5_300k_tc_oneshot.json 2> /dev/null
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c4_tc_pipe2.json 2> gpurun_out/bench_c4_tc_pipe2.err
python - <<'PY'
import json
for f in ("bench_c5_300k_tc","bench_c5_300k_tc_oneshot","bench_c4_tc_pipe2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d["roofline"]
        print(f, "value %.4g ms %.2f e2e %.4g ms %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["e2e"].get("pipeline_ranges"), d["config"]["variant"], d["config"]["edges"], "filter_ms", r["launch_ms"], "exec TF", r.get("executed_tflops"), r["other_kernels_ms"])
    except Exception as e: print(f, "FAILED", e)
PY
