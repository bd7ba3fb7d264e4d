#!/bin/bash
# Round 2, session 1: TMEM read-out probe, ncu capture of k_exact_queue (K3), baseline bench line of this box.
mkdir -p gpurun_out
timeout 300 ./tools/tmem_probe > gpurun_out/r02_tmem_probe.json 2> gpurun_out/r02_tmem_probe.err; echo "probe rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c4_base.json 2> gpurun_out/r02_bench_c4_base.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_exact_queue -s 3 -c 1 -o gpurun_out/r02_prof_exact_queue \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_exact_queue.log 2>&1; echo "ncu rc=$?"
head -c 600 gpurun_out/r02_bench_c4_base.json; echo
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_tmem_probe.json"))
for r in d["results"]: print(r)
PY
