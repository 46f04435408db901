#!/bin/bash
mkdir -p gpurun_out
for r in 0 1 2 3 4 5 6 7; do
timeout 300 python bench.py --partition-of 8 --partition-rank $r --steps 24 --warmup 4 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/bench_p8_r${r}.json 2> gpurun_out/bench_p8_r${r}.err
python - <<PY
import json
for l in open('gpurun_out/bench_p8_r${r}.json'):
    if l.startswith('{'):
        d = json.loads(l); print('rank $r of 8', d['ms_per_step'], d.get('pass_ms'))
PY
done
