#!/bin/bash
# 8-GPU call: the bench line at N = 8 exactly as the driver launches it
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/bench_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_n8.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('n8', d['value'], d['ms_per_step'], json.dumps(d.get('e2e')))
        print(json.dumps(d.get('pass_ms')), json.dumps(d.get('clocks')))
        for k, v in d.get('secondary', {}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('pass_ms'))
PY
