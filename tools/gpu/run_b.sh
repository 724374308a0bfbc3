#!/bin/bash
# GPU call B (N GPUs): default bench under torchrun with the split sub-records.
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/b_bench_n$N.json 2> gpurun_out/b_bench_n$N.err; tail -6 gpurun_out/b_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/b_bench_n$N.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('c2', d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), 'e2e', d['e2e'] and (round(d['e2e']['value']), d['e2e'].get('frac_of_copy_ceiling'), d['e2e']['copy_ceiling']['node_total_gbs']))
    for k,v in d.get('splits',{}).items():
        print(k, v.get('error') or (round(v['value']), round(v['ms_per_step'],4), v['scaling'], v.get('e2e') and round(v['e2e']['value'])))
except Exception as e:
    print('bench parse failed', e)
PY
