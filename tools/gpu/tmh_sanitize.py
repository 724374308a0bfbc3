"""Small fftfilt_tmh_kernel runs for compute-sanitizer (memcheck / racecheck / synccheck): variants 40 / 41 / 42, ragged sizes, carried
history, tap partitions (accumulating launches), fused decimation by a non-polyphase factor."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from oracle import oracle as O
import rustradio_b200 as R

for variant in ("42", "41", "40"):
    os.environ["RRC_FFTFILT_VARIANT"] = variant
    for ntaps, n in ((4097, 100_003), (193, 40_000), (16385, 90_000)):
        taps = O.low_pass_n(1.0, 0.05, ntaps).astype(np.complex64) * (1 + 0.3j)
        x = O.synth_c32(11, 0, n)
        truth = O.conv_full_f64_fft(x, taps, n)
        f = R.FftFilt(taps)
        cut = n // 2 + 3
        got = []
        for lo, hi in ((0, cut), (cut, n)):
            din = R.DeviceBuffer.from_numpy(np.ascontiguousarray(x[lo:hi]))
            dout = R.DeviceBuffer((hi - lo) * 8)
            f.run(din, hi - lo, dout)
            got.append(dout.download(np.complex64, hi - lo))
        e = O.rel_rms(np.concatenate(got), truth)
        print(variant, ntaps, n, f"{e:.2e}")
        assert e < 1e-5
    taps = O.low_pass_n(1.0, 0.05, 4097).astype(np.complex64)
    x = O.synth_c32(12, 0, 80_000)
    f = R.FftFilt(taps)
    din = R.DeviceBuffer.from_numpy(x)
    dout = R.DeviceBuffer(len(x) * 8)
    cnt = f.decim_run(din, len(x), 1000, 7, dout)
    e = O.rel_rms(dout.download(np.complex64, cnt), O.conv_full_f64_fft(x, taps, len(x))[7::1000])
    print(variant, "decim 1000", f"{e:.2e}")
    assert e < 1e-5
print("done")
