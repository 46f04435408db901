#!/bin/bash
mkdir -p gpurun_out
for p in 8 4; do for v in 2 3 4; do
timeout 300 python bench.py --partition-of $p --views-in-flight $v --steps 24 --warmup 4 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/bench_p${p}_v${v}.json 2> gpurun_out/bench_p${p}_v${v}.err
python - <<PY
import json
for l in open('gpurun_out/bench_p${p}_v${v}.json'):
    if l.startswith('{'):
        d = json.loads(l); print('partition-of $p inflight $v', d['ms_per_step'])
PY
done; done
for v in 3 4; do
timeout 300 python bench.py --views-in-flight $v --steps 12 --warmup 4 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/bench_p1_v${v}.json 2> gpurun_out/bench_p1_v${v}.err
python - <<PY
import json
for l in open('gpurun_out/bench_p1_v${v}.json'):
    if l.startswith('{'):
        d = json.loads(l); print('N=1 inflight $v', d['ms_per_step'])
PY
done
