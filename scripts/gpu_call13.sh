#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "prb or deferred or golden" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 600 python scripts/prb_timing.py > gpurun_out/prb_timing_queue.log 2>&1; echo "prb rc=$?"; grep "_ms" gpurun_out/prb_timing_queue.log | tail -4
SDFRT_SITEQ=0 timeout 600 python scripts/prb_timing.py > gpurun_out/prb_timing_noqueue.log 2>&1; echo "prb rc=$?"; grep "_ms" gpurun_out/prb_timing_noqueue.log | tail -4
timeout 600 python bench.py --workload c5_prb_blobs256_512x512_128spp --steps 3 --warmup 1 --no-secondary --no-e2e --no-grad-check > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"; tail -2 gpurun_out/bench_c5.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_c5.json'):
    if l.startswith('{'):
        d = json.loads(l); print('c5', d['value'], d['ms_per_step'])
PY
