"""Small fir_tc5_kernel runs for compute-sanitizer (memcheck / racecheck / synccheck): ragged, multi-channel, both tile heights."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
os.environ["RRC_FIR_TCGEN05"] = "2"
from oracle import oracle as O
import rustradio_b200 as R

for rows in ("32", "64"):
    os.environ["RRC_FIR_TC5_NR"] = rows
    for T, n, nchan in ((64, 3 * 8192 + 77, 1), (33, 5000, 3), (65, 8192 * 2 + 64, 2)):
        taps = O.low_pass_n(1.0, 0.2, T).astype(np.complex64)
        f = R.Fir(taps)
        stride = n + 1
        xs = np.zeros((nchan, stride), np.complex64)
        for c in range(nchan):
            xs[c, :n] = O.synth_c32(7 + c, 0, n)
        out_n = f.out_count(n)
        din = R.DeviceBuffer.from_numpy(xs)
        ostride = out_n + 1 - (out_n % 2)
        dy = R.DeviceBuffer(nchan * ostride * 8)
        f.run_batch(din, stride, out_n - 1 + T, dy, ostride, out_n, nchan)
        y = dy.download(np.complex64, nchan * ostride).reshape(nchan, ostride)[:, :out_n]
        e = max(O.rel_rms(y[c], O.fir(xs[c, :n], taps, 1, f64=True)) for c in range(nchan))
        print(rows, T, n, nchan, f.kernel_name[:14], f"{e:.2e}")
        assert e < 2e-6
print("done")
