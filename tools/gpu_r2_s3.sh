#!/bin/bash
# Round 2, session 3: ncu --set full of the restructured tcgen05 filter, CG=1 and CG=2.
mkdir -p gpurun_out
for cg in 1 2; do
SCEMA_TC_CG=$cg timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_filter_tc -s 3 -c 1 -o gpurun_out/r02_prof_filter_tc_cg$cg \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_filter_tc_cg$cg.log 2>&1; echo "ncu cg$cg rc=$?"
done
