"""Small fftfilt_poly_kernel runs for compute-sanitizer (memcheck / racecheck / synccheck): every cluster width, pair gathers and
the unaligned fallback, history carried across calls, several iterations per cluster (so that the scratch slots are reused)."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from oracle import oracle as O
import rustradio_b200 as R

small = os.environ.get("POLY_SANITIZE_SMALL") == "1"
for C in ("4", "2", "1"):
    os.environ["RRC_FFTFILT_POLY_C"] = C
    os.environ["RRC_FFTFILT_POLY_GROUPS"] = "2"          # 2 clusters: every cluster iterates several blocks
    for ntaps, n, deci, skip in ((16385, 160_000 if small else 1_300_000, 8, 0), (4097, 90_000, 8, 3), (301, 70_000, 4, 1)):
        taps = O.low_pass_n(1.0, 0.02, ntaps).astype(np.complex64) * (1 - 0.2j)
        x = O.synth_c32(9, 0, n)
        truth = O.conv_full_f64_fft(x, taps, n)
        f = R.FftFilt(taps)
        cut = n // 2 + 1
        got = []
        for lo, hi in ((0, cut), (cut, n)):
            sk = (skip - lo) % deci if lo else skip
            din = R.DeviceBuffer.from_numpy(np.ascontiguousarray(x[lo:hi]))
            dout = R.DeviceBuffer(((hi - lo) // deci + 2) * 8)
            cnt = f.decim_run(din, hi - lo, deci, sk, dout)
            got.append(dout.download(np.complex64, cnt))
        got = np.concatenate(got)
        want = truth[skip::deci]
        assert len(got) == len(want), (len(got), len(want))
        e = O.rel_rms(got, want)
        print(C, ntaps, n, deci, skip, f"{e:.2e}")
        assert e < 1e-5
print("done")
