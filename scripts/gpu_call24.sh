#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_p8.csv python bench.py --partition-of 8 --steps 2 --warmup 1 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/ncu_p8.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --partition-of 8 --steps 40 --warmup 4 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/bench_p8.json 2> gpurun_out/bench_p8.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_p8.json'):
    if l.startswith('{'):
        d = json.loads(l); print('p8', d['ms_per_step'], json.dumps(d['roofline'].get('launch_ms_union_per_step')))
PY
