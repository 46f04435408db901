#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "not tma_box" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
for v in 1 0; do
SDFRT_DEFER_SIDE=$v timeout 300 python bench.py --partition-of 8 --steps 40 --warmup 4 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/bench_p8_side$v.json 2> gpurun_out/bench_p8_side$v.err
SDFRT_DEFER_SIDE=$v timeout 300 python bench.py --steps 12 --warmup 3 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/bench_p1_side$v.json 2> gpurun_out/bench_p1_side$v.err
python - <<PY
import json
for f in ('gpurun_out/bench_p8_side$v.json', 'gpurun_out/bench_p1_side$v.json'):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print(f, round(d['ms_per_step'], 3))
PY
done
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/bench.json'):
    if l.startswith('{'):
        d = json.loads(l); print('default bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['grad_check']['sdf']['rel_l2'])
PY
