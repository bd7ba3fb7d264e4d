#!/bin/bash
# Round 2, session 16: k_resample_pair on the GPU — the whole GPU suite with the pair kernel, K1 alone in five variants
# (first-generation kernel, pair with 5 / 4 / 6 resident CTAs per SM, pair with 12 warps per SM), ncu --set full of the pair kernel.
mkdir -p gpurun_out
export SCEMA_K1_KERNEL=pair
timeout 250 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02_pytest_pair.log 2>&1; echo "pytest(pair) rc=$?"
tail -4 gpurun_out/r02_pytest_pair.log
probe() {  # name, kernel, lib, wps
    SCEMA_K1_KERNEL=$2 SCEMA_LIB=$3 SCEMA_K1_WPS=$4 timeout 60 python tools/k1_probe.py > gpurun_out/r02_k1_probe_$1.json 2> gpurun_out/r02_k1_probe_$1.err
    echo "== $1 rc=$?"; cat gpurun_out/r02_k1_probe_$1.err
}
probe stream stream "" ""
probe pair_m5 pair "" ""
probe pair_m4 pair $PWD/scema_b200/libscema_hist_m4.so ""
probe pair_m6 pair $PWD/scema_b200/libscema_hist_m6.so ""
probe pair_m5_w12 pair "" 12
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_resample_pair -s 6 -c 2 -o gpurun_out/r02_prof_resample_pair_c3 \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --verify off > gpurun_out/r02_ncu_resample_pair.log 2>&1; echo "ncu rc=$?"
