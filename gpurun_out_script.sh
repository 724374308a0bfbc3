mkdir -p gpurun_out
for cfg in c1f c1d2; do
timeout 300 python bench.py --config $cfg --steps 20 --warmup 3 > gpurun_out/bench_${cfg}_v11.json 2> gpurun_out/bench_${cfg}_v11.err; tail -3 gpurun_out/bench_${cfg}_v11.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_${cfg}_v11.json').read().strip().splitlines()[-1]); print('$cfg', round(d['value']), d['ms_per_step'], round(d['roofline']['frac'],3), d['roofline']['kernel'][:34], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline'].get('fp32',{}).get('frac'))"
env RRC_FIR_TENSOR=0 timeout 300 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${cfg}_fp32.json 2> gpurun_out/bench_${cfg}_fp32.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_${cfg}_fp32.json').read().strip().splitlines()[-1]); print('$cfg fp32', round(d['value']), d['ms_per_step'], round(d['roofline']['frac'],3), d['roofline']['kernel'][:34])"
done
timeout 200 python bench.py --config c1f --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-400
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fir_tcf_kernel -s 3 -c 1 -f -o gpurun_out/c1f_v11 python bench.py --config c1f --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_c1f.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fir_tc1_kernel -s 3 -c 1 -f -o gpurun_out/c1d2_v11 python bench.py --config c1d2 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_c1d2.log 2>&1
