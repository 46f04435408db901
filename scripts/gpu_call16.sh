#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "prb or deferred or golden or parity or convolution" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 600 python scripts/conv_timing.py > gpurun_out/conv_timing.log 2>&1; echo "conv rc=$?"; tail -7 gpurun_out/conv_timing.log
timeout 600 python scripts/prb_timing.py > gpurun_out/prb_timing.log 2>&1; echo "prb rc=$?"; grep "_ms" gpurun_out/prb_timing.log | tail -4
for v in 1 0; do
SDFRT_SHADE_ROWS8=$v timeout 600 python bench.py --steps 12 --warmup 3 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/bench_c3_shade$v.json 2> gpurun_out/bench_c3_shade$v.err; echo "c3 shade_rows8=$v rc=$?"
python - <<PY
import json
for l in open('gpurun_out/bench_c3_shade$v.json'):
    if l.startswith('{'):
        d = json.loads(l); print('c3 shade_rows8=$v', d['value'], d['ms_per_step'])
PY
done
timeout 600 python bench.py --workload c5_prb_blobs256_512x512_128spp --steps 3 --warmup 1 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_c5.json'):
    if l.startswith('{'):
        d = json.loads(l); print('c5', d['value'], d['ms_per_step'])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_prb_backward|k_scatter_sites" -c 2 -o gpurun_out/c5_prb_bwd_r2 python bench.py --workload c5_prb_blobs256_512x512_128spp --steps 1 --warmup 0 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/ncu_c5_full.log 2>&1; echo "ncu full rc=$?"
python profiles/ncu_summary.py gpurun_out/c5_prb_bwd_r2.ncu-rep > gpurun_out/r2_c5_prb_bwd_queue.md 2> gpurun_out/ncu_summary.err; echo "summary rc=$?"
rm -f gpurun_out/*.ncu-rep
