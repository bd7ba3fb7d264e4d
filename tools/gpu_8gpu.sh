#!/bin/bash
# 8-GPU session: sharded parity check, config 4 at 8 GPUs, config 5 at full size (4M x 300) streamed.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/smi8.txt
timeout 600 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -3
run() { # name, extra args
  timeout $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 8 $2 > gpurun_out/$1.json 2> gpurun_out/$1.err
  python - gpurun_out/$1.json $1 <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print("%-14s value=%.4g ms/step=%.2f e2e=%.4g filter_ms=%.2f TF=%.2f frac=%.3f edges=%d other=%s" % (sys.argv[2], d["value"], d["ms_per_step"], d["e2e"]["value"], r["launch_ms"], r["achieved"], r["frac"], d["config"]["edges"], {k: round(v,3) for k,v in r["other_kernels_ms"].items()}))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
  grep -v "^W1\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/$1.err | tail -3
}
run bench_c4_g8 "--steps 5 --warmup 3" 300 29561
run bench_c5_g8 "--workload c5 --steps 1 --warmup 3" 900 29562
