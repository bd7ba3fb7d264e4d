#!/bin/bash
# Round 2, session 17: K1 tuning grid (kernel x warps per SM x memory flags) in one process per build of k_resample_pair
# (5 / 4 / 6 resident CTAs per SM).
mkdir -p gpurun_out
for v in m5 m4 m6; do
    lib=""; [ $v != m5 ] && lib=$PWD/scema_b200/libscema_hist_$v.so
    SCEMA_LIB=$lib timeout 90 python tools/k1_probe.py --scan > gpurun_out/r02_k1_scan_$v.json 2> gpurun_out/r02_k1_scan_$v.txt
    echo "== $v rc=$?"; grep "^best\|stream" gpurun_out/r02_k1_scan_$v.txt
done
