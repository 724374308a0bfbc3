timeout 300 python bench.py --config c2 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 normal', round(d['ms_per_step'],4))"
cp rustradio_b200/librustradio_cuda.so /tmp/keep.so; cp gpurun_exp_nopowers.so rustradio_b200/librustradio_cuda.so
timeout 300 python bench.py --config c2 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 no-powers what-if', round(d['ms_per_step'],4))"
