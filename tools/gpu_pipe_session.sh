#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x --durations=5 > gpurun_out/pytest_tc.log 2>&1
echo "exit $?" >> gpurun_out/pytest_tc.log; tail -n 25 gpurun_out/pytest_tc.log
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 --deselect tests/test_gpu_tc.py > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log; tail -n 12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c4_tc_pipe.json 2> gpurun_out/bench_c4_tc_pipe.err
echo "exit $?" >> gpurun_out/bench_c4_tc_pipe.err; cat gpurun_out/bench_c4_tc_pipe.json; tail -3 gpurun_out/bench_c4_tc_pipe.err
SCEMA_PIPELINE=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c4_tc_nopipe.json 2> /dev/null
python - <<'PY'
import json
for f in ("bench_c4_tc_pipe","bench_c4_tc_nopipe"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, "value %.4g ms %.2f e2e %.4g ms %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["e2e"].get("pipeline_ranges"), d["roofline"]["other_kernels_ms"])
    except Exception as e: print(f, "FAILED", e)
PY
