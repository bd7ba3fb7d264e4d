#!/bin/bash
# A/B of library builds on one box: SCEMA_LIB selects the .so (scema_b200/libscema_hist_<tag>.so)
mkdir -p gpurun_out
for lib in "$@"; do
  L=$PWD/scema_b200/libscema_hist.so; [ $lib != new ] && L=$PWD/scema_b200/libscema_hist_$lib.so
  SCEMA_LIB=$L SCEMA_TC_SLICES=1 timeout 300 python tools/tc_probe.py 1000 1000000 2>&1 | grep "big n" | sed -n 2p | sed "s/^/$lib s1: /"
  SCEMA_LIB=$L SCEMA_TC_SLICES=2 timeout 300 python tools/tc_probe.py 1000 1000000 2>&1 | grep "big n" | sed -n 2p | sed "s/^/$lib s2: /"
done
