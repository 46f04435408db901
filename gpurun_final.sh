mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python -c "
import json
d=json.load(open('gpurun_out/bench_default.json')); r=d['roofline']
print(round(d['value'],1), round(d['ms_per_step'],2), d['e2e'], d['clocks'], d['gpu_launches'], round(r['frac'],3), r['traffic'], d['cpu_baseline'])
print(json.load(open('gpurun_out/bench_reference.json'))['value'])"
tail -2 gpurun_out/bench_default.err
