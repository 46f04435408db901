#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "partition or band or counters" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
for rows in 2 1; do for r in 0 1 2 3 4 5 6 7; do
timeout 300 python bench.py --partition-of 8 --partition-rank $r --partition-rows $rows --steps 24 --warmup 4 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/bench_p8_rows${rows}_r${r}.json 2> gpurun_out/bench_p8_rows${rows}_r${r}.err
python - <<PY
import json
for l in open('gpurun_out/bench_p8_rows${rows}_r${r}.json'):
    if l.startswith('{'):
        d = json.loads(l); print('rows $rows rank $r of 8', round(d['ms_per_step'], 3))
PY
done; done
