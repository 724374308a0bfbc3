#!/bin/bash
mkdir -p gpurun_out
for c in c2 c1 c4 c3 c5; do
timeout 300 python bench.py --config $c --steps 5 --warmup 3 --headline-only --no-cpu --sustain 0 > gpurun_out/o_tmp.json 2> gpurun_out/o_e2e.err; python -c "import json;d=json.load(open('gpurun_out/o_tmp.json'));e=d['e2e'];print('$c adaptive chunk e2e', round(e['value']), 'Msps', round(e['ms_per_step'],2),'ms  ceiling', round(e['copy_ceiling']['ms_per_step'],2), 'frac', round(e['frac_of_copy_ceiling'],3))"
done 2>&1 | tee gpurun_out/o_e2e_adaptive.txt
timeout 900 python -m pytest tests -q -m gpu -x -k "host or e2e or pipe or run_host" 2>&1 | tail -3
