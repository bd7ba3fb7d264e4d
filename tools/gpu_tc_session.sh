#!/bin/bash
# One single-GPU session for the tcgen05 filter: probe (layout, error model, timings per kernel flavour), the GPU
# suite, both bench arms, launch list and ncu --set full captures of k_filter_tc. Everything lands in gpurun_out/.
mkdir -p gpurun_out
for combo in "2 1" "2 2" "1 1"; do
  set -- $combo
  SCEMA_TC_CG=$1 SCEMA_TC_SLICES=$2 timeout 300 python tools/tc_probe.py 1000 1000000 > gpurun_out/tc_probe_cg$1_s$2.log 2>&1
  echo "exit $?" >> gpurun_out/tc_probe_cg$1_s$2.log
done
tail -n 4 gpurun_out/tc_probe_cg*_s*.log
timeout 900 python -m pytest tests/test_gpu_tc.py -q --durations=8 > gpurun_out/pytest_tc.log 2>&1
echo "exit $?" >> gpurun_out/pytest_tc.log; tail -n 5 gpurun_out/pytest_tc.log
timeout 1500 python -m pytest tests -m gpu -q --durations=12 --deselect tests/test_gpu_tc.py > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log; tail -n 5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_c4_tc.json 2> gpurun_out/bench_c4_tc.err
echo "exit $?" >> gpurun_out/bench_c4_tc.err; cat gpurun_out/bench_c4_tc.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
SCEMA_TC_SLICES=2 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c4_tc_s2.json 2> gpurun_out/bench_c4_tc_s2.err
timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_tc.json 2> gpurun_out/bench_c3_tc.err
# launch list of the default bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c4_tc.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_filter_tc -s 3 -c 1 -o gpurun_out/prof_filter_tc_s1 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_filter_tc_s1.log 2>&1
SCEMA_TC_SLICES=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_filter_tc -s 3 -c 1 -o gpurun_out/prof_filter_tc_s2 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_filter_tc_s2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_tc_prep -s 3 -c 1 -o gpurun_out/prof_tc_prep \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tc_prep.log 2>&1
ls -la gpurun_out | tail -20
