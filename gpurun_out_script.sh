mkdir -p gpurun_out
RRC_FFTFILT_VARIANT=35 timeout 600 python -m pytest tests -m gpu -x -q -k "fftfilt or FftFilt or fft" 2>&1 | tail -3
for v in 32 35; do
  RRC_FFTFILT_VARIANT=$v timeout 300 python bench.py --config c2 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/v_$v.json 2> gpurun_out/v_$v.err
  python -c "import json;d=json.load(open('gpurun_out/v_$v.json'));print('variant $v', round(d['ms_per_step'],4))"
done
