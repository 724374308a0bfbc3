#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_egress.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/o_pytest.txt
python - <<'PY'
import numpy as np, time, torch
import rustradio_b200 as R
n = 1 << 28
din = R.DeviceBuffer(n * 8); R.synth_f32(din, 5, 0, 2 * n)
dout = R.DeviceBuffer(2 * n)
for _ in range(3): R.rtlsdr_encode(din, n, dout)
R.device_sync()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s = torch.cuda.current_stream()
ev0.record(s)
for _ in range(10): R.rtlsdr_encode(din, n, dout, 0, s.cuda_stream)
ev1.record(s); ev1.synchronize()
ms = ev0.elapsed_time(ev1) / 10
print(f"rtlsdr_encode 2^28 samples: {ms:.3f} ms, {n * 10 / ms / 1e6:.0f} GB/s algorithmic (8 B in + 2 B out per sample)")
PY
