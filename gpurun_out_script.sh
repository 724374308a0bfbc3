mkdir -p gpurun_out
for cfg in c3 c4 c5 f1 f2 h1 e1 e2 e3 e4 a12 c3u8 c5u8; do
timeout 400 python bench.py --config $cfg --steps 20 --warmup 3 --cpu-budget 6 > gpurun_out/bench_${cfg}_v11.json 2> gpurun_out/bench_${cfg}_v11.err
python -c "
import json
try:
    d=json.loads(open('gpurun_out/bench_${cfg}_v11.json').read().strip().splitlines()[-1]); print('$cfg', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), (round(d['e2e']['value']) if d.get('e2e') else None), (round(d['cpu_baseline']['value'],1) if d.get('cpu_baseline') else None))
except Exception as e: print('$cfg failed', e)
"
done
