#!/bin/bash
# Round 2, session 19: ring depth 16 (15 steps of prefetch) against the default 8, K1 alone; K1 parity tests with the deeper ring.
mkdir -p gpurun_out
timeout 40 python tools/k1_probe.py > gpurun_out/r02_k1_probe_d8.json 2> gpurun_out/r02_k1_probe_d8.err; echo "d8 rc=$?"; cat gpurun_out/r02_k1_probe_d8.err
SCEMA_LIB=$PWD/scema_b200/libscema_hist_d16.so timeout 40 python tools/k1_probe.py > gpurun_out/r02_k1_probe_d16.json 2> gpurun_out/r02_k1_probe_d16.err; echo "d16 rc=$?"; cat gpurun_out/r02_k1_probe_d16.err
SCEMA_LIB=$PWD/scema_b200/libscema_hist_d16.so timeout 60 python -m pytest tests/test_gpu_parity.py -q -x -k "k1 or history_store or empty_batch" > gpurun_out/r02_pytest_k1_d16.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_k1_d16.log
