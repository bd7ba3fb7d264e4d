#!/bin/bash
# One single-GPU session on the B200 box (run through gpurun): parity suite, smoke, the default bench
# with both arms, config 3 / config-5-shape benches, launch list and full ncu captures of K2 and K1,
# compute-sanitizer on the small parity cases. Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 2>&1 | tail -14 > gpurun_out/pytest_gpu.txt; tail -14 gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d.get("roofline") or {}
    print("%-22s value=%.4g ms/step=%.2f e2e=%.4g" % (sys.argv[2], d["value"], d["ms_per_step"], d["e2e"]["value"]), "filter_ms=%s TF=%s frac=%s other=%s" % (r.get("launch_ms"), r.get("achieved"), r.get("frac"), r.get("other_kernels_ms")))
    if d.get("cpu_baseline"): print("   cpu:", {k: v for k, v in d["cpu_baseline"].items() if k != "sample"})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; show gpurun_out/bench_ref.json "reference arm"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; show gpurun_out/bench_c4.json "default (c4)"
timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; show gpurun_out/bench_c3.json "c3"
timeout 600 python bench.py --workload c5 --histories 300000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_300k.json 2> gpurun_out/bench_c5_300k.err; show gpurun_out/bench_c5_300k.json "c5 shape, 300k"
# launch list (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_filter_ws -s 3 -c 1 -o gpurun_out/prof_filter \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_filter.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resample_stream -s 6 -c 2 -o gpurun_out/prof_resample \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_resample.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or special or empty or store or stream_concat" > gpurun_out/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k2_golden or k1_golden" > gpurun_out/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"
