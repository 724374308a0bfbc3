mkdir -p gpurun_out
for cfg in c1 c2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $cfg --steps 20 --warmup 3 --no-cpu > gpurun_out/scale2_${cfg}.json 2> gpurun_out/scale2_${cfg}.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/scale2_${cfg}.json') if l.startswith('{')][-1]); print('$cfg N=2', round(d['value']), d['ms_per_step'], d['n_gpus'], d['scaling'], round(d['e2e']['value']) if d['e2e'] else None)"
done
