mkdir -p gpurun_out
for v in 32 36 32 36; do
RRC_FFTFILT_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 8 > gpurun_out/b.json 2> gpurun_out/b.err; tail -2 gpurun_out/b.err
python - <<PY
import json
d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print('variant $v', round(d['value']), d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
PY
done
