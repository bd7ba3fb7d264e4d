#!/bin/bash
# Round 2, session 9: K1 with guard-free interior blocks: parity + probe.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "k1 or store or config3 or ids" 2>&1 | tail -4
timeout 600 python tools/k1_probe.py 200000 500 2>&1 >gpurun_out/r02_k1_probe_v2.json | tail -4
