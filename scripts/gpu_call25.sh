#!/bin/bash
mkdir -p gpurun_out
for c in 3 2 4; do
timeout 600 python bench.py --e2e-contexts $c --steps 8 --warmup 3 --no-secondary --no-grad-check --no-cpu-baseline > gpurun_out/bench_e2e_c$c.json 2> gpurun_out/bench_e2e_c$c.err; echo "rc=$?"
python - <<PY
import json
for l in open('gpurun_out/bench_e2e_c$c.json'):
    if l.startswith('{'):
        d = json.loads(l); print('e2e contexts $c', d['value'], json.dumps(d['e2e']))
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_p8.csv python bench.py --partition-of 8 --steps 2 --warmup 1 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/ncu_p8.log 2>&1; echo "ncu rc=$?"
