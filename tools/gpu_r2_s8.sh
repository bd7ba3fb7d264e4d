#!/bin/bash
# Round 2, session 8: K1 alone, occupancy variants.
mkdir -p gpurun_out
for cfg in "6 24" "8 32" "6 20"; do
  set -- $cfg
  echo "MINB=$1 WPS=$2"
  SCEMA_K1_MINB=$1 SCEMA_K1_WPS=$2 timeout 600 python tools/k1_probe.py 200000 500 2>&1 >gpurun_out/r02_k1_probe_minb$1_wps$2.json | tail -4
done
