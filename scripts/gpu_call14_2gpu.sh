#!/bin/bash
# 2-GPU call: NCCL parity test and the bench line at N = 2 exactly as the driver launches it (e2e leg with double-buffered copies)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_n2.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('n2', d['value'], d['ms_per_step'], json.dumps(d.get('e2e')))
        for k, v in d.get('secondary', {}).items(): print(k, v.get('value'), v.get('ms_per_step'))
PY
