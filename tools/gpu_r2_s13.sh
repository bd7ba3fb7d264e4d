#!/bin/bash
# Round 2, session 13: ncu --set full of the final tcgen05 filter (default cta_group::2), compute-sanitizer on small tcgen05
# cases, the new select_rows test.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "select_rows or store" 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_filter_tc -s 3 -c 1 -o gpurun_out/r02_prof_filter_tc_final \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --verify off > gpurun_out/r02_ncu_filter_tc_final.log 2>&1; echo "ncu rc=$?"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "accumulators and (clusters or k18) or production or chosen_up_front or (wide_rows and 50) or norm_band" > gpurun_out/r02_sanitizer_memcheck_tc.txt 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/r02_sanitizer_memcheck_tc.txt
