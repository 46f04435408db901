#!/bin/bash
mkdir -p gpurun_out
for b in 0 24; do
if [ $b = 0 ]; then unset SDFRT_BATCH_LOG2; else export SDFRT_BATCH_LOG2=$b; fi
timeout 300 python bench.py --partition-of 8 --steps 40 --warmup 4 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/bench_p8_b$b.json 2> gpurun_out/bench_p8_b$b.err
python - <<PY
import json
for l in open('gpurun_out/bench_p8_b$b.json'):
    if l.startswith('{'):
        d = json.loads(l); print('partition-of 8 batch_log2=$b', round(d['ms_per_step'], 3))
PY
done
unset SDFRT_BATCH_LOG2
for o in 2 3; do
SDFRT_PRB_OCC=$o timeout 600 python bench.py --workload c5_prb_blobs256_512x512_128spp --steps 4 --warmup 1 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/bench_c5_occ$o.json 2> gpurun_out/bench_c5_occ$o.err
python - <<PY
import json
for l in open('gpurun_out/bench_c5_occ$o.json'):
    if l.startswith('{'):
        d = json.loads(l); print('c5 occ=$o', d['value'], d['ms_per_step'])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_final_launches_c3.csv python bench.py --steps 1 --warmup 1 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2_final_launches_c3.csv
