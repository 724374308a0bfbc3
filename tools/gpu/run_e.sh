#!/bin/bash
# GPU call E: ncu --set full on the packed (37) and scalar TMA (36) FftFilter kernels, 2^27 samples.
mkdir -p gpurun_out
for v in 37 36; do
  k=$([ $v = 37 ] && echo fftfilt_pk_kernel || echo fftfilt_tma_kernel)
  RRC_FFTFILT_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/e_c2_v$v \
     python bench.py --config c2 --n 134217728 --steps 2 --warmup 3 --headline-only --no-e2e --no-cpu --sustain 0 > gpurun_out/e_ncu_v$v.log 2>&1
  echo "ncu v$v rc=$?"
done
ls -la gpurun_out/*.ncu-rep
