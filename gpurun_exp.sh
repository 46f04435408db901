timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for sl in 4 8; do
  echo -n "SLOTS $sl: "; SDFRT_SLOTS=$sl timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(round(d['value'],1), round(d['ms_per_step'],2), {k: round(v,2) for k,v in r['pass_ms'].items()})"
done
echo -n "C4 512^3: "; timeout 400 python bench.py --workload c4_blobs512_512x512_64spp --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(round(d['value'],1), round(d['ms_per_step'],2), {k: round(v,2) for k,v in r['pass_ms'].items()}, round(r['frac'],3), d['e2e']['value'])"
