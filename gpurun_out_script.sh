mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/scale2_c2.json 2> gpurun_out/scale2_c2.err; tail -3 gpurun_out/scale2_c2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --shard time --no-e2e --no-cpu > gpurun_out/scale2_c2_time.json 2> gpurun_out/scale2_c2_time.err; tail -3 gpurun_out/scale2_c2_time.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/scale2_ref.json 2> gpurun_out/scale2_ref.err; tail -3 gpurun_out/scale2_ref.err
python - <<PY
import json
for f in ('scale2_c2','scale2_c2_time','scale2_ref'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, d.get('n_gpus'), round(d['value']), d.get('scaling'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))
    except Exception as e: print(f, 'failed', e)
PY
