"""Diagnostics of fir_tc5_kernel (tcgen05 FIR) on the GPU box: delta taps locate addressing errors, then parity."""
import os
import sys
import numpy as np

sys.path.insert(0, os.getcwd())
os.environ.setdefault("RRC_FIR_TCGEN05", "1")
from oracle import oracle as O
import rustradio_b200 as R


def report(tag, y, want):
    if len(y) != len(want):
        print(f"{tag}: LENGTH {len(y)} != {len(want)}")
        return False
    if len(y) == 0:
        print(f"{tag}: empty, ok")
        return True
    bad = ~np.isclose(y, want, rtol=1e-4, atol=1e-5 * float(np.abs(want).max() + 1e-30))
    e = O.rel_rms(y, want)
    print(f"{tag}: rel_rms {e:.3e} bad {int(bad.sum())}/{len(y)}")
    if bad.any():
        idx = np.nonzero(bad)[0]
        print("   first bad:", idx[:12].tolist(), " o%64 hist:", np.bincount(idx % 64, minlength=64).tolist())
        print("   (o//64)%8 hist:", np.bincount((idx // 64) % 8, minlength=8).tolist(), " tile hist:", np.bincount(idx // 8192)[:6].tolist())
    return not bad.any()


def delta_map(y, x, o_list):
    """For delta taps every output equals some input sample: say which."""
    out = []
    for o in o_list:
        d = np.abs(x - y[o])
        k = int(np.argmin(d))
        out.append((o, k - o if d[k] < 1e-3 * abs(x[k]) else None))
    return out


for bo in (0,):
    os.environ["RRC_FIR_TC5_BASE_OFF"] = str(bo)
    print(f"===== base_off {bo}")
    n = 3 * 8192 + 1234
    x = O.synth_c32(5, 0, n)
    ok_all = True
    for T, j in ((65, 64), (65, 0), (65, 1), (65, 9), (65, 63), (33, 20)):
        w = np.zeros(T, np.float32); w[j] = 1.0            # reversed taps: y[o] = x[o + j]
        taps = w[::-1].astype(np.complex64)
        f = R.Fir(taps)
        if "tc5" not in f.kernel_name:
            print("NOT tc5:", f.kernel_name); break
        y = f.filter(x)
        want = x[j:j + len(y)]
        ok = report(f"delta T={T} j={j}", y, want)
        ok_all &= ok
        if not ok:
            print("   map (o, src - o):", delta_map(y, x, [0, 1, 2, 8, 63, 64, 65, 127, 128, 512, 8191, 8192]))
    for T, nn in ((64, 100_000), (65, 8192 + 64), (17, 50_001), (40, 7), (64, 64), (64, 8192 * 5 + 63), (33, 8192 * 2 + 32)):
        xx = O.synth_c32(6, 0, nn)
        taps = O.low_pass_n(1.0, 0.1, T).astype(np.complex64)
        f = R.Fir(taps)
        y = f.filter(xx)
        ok_all &= report(f"lowpass T={T} n={nn} [{f.kernel_name[:14]}]", y, O.fir(xx, taps, 1, f64=True))
    print("base_off", bo, "ALL OK" if ok_all else "FAILED")
