#!/bin/bash
# Round 2, session 18: validation of the K1 defaults — whole GPU suite, smoke with audit, K1 alone, bench config 4, ncu of K1 on config 4.
mkdir -p gpurun_out
timeout 250 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02_pytest_k1pair_final.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest_k1pair_final.log
timeout 60 python __graft_entry__.py smoke > gpurun_out/r02_smoke_k1pair.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke_k1pair.log
timeout 60 python tools/k1_probe.py > gpurun_out/r02_k1_probe_final.json 2> gpurun_out/r02_k1_probe_final.err; echo "probe rc=$?"; cat gpurun_out/r02_k1_probe_final.err
timeout 120 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_c4_k1pair.json 2> gpurun_out/r02_bench_c4_k1pair.err; echo "bench rc=$?"; head -c 1500 gpurun_out/r02_bench_c4_k1pair.json
timeout 60 ncu --set full --clock-control none --import-source on -k regex:k_resample_pair -s 3 -c 1 -o gpurun_out/r02_prof_resample_pair_c4 \
    python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline --verify off > gpurun_out/r02_ncu_resample_pair_c4.log 2>&1; echo "ncu rc=$?"
