"""CPU emulation of the FftFilter CUDA kernel (tests/emul/fftfilt_emul.cu): the
kernel's phase functions are __host__ __device__, so the exact index math,
shared-memory layout and twiddle logic run here thread-by-thread on the CPU and
are checked against the oracle.  This is a test of the kernel SOURCE without a
GPU; it is not a product path (nothing in rustradio_b200 can reach it)."""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O

HERE = Path(__file__).resolve().parent / "emul"
SO = HERE / "_fftfilt_emul.so"


@pytest.fixture(scope="module")
def emul():
    if shutil.which("nvcc") is None and not Path("/usr/local/cuda/bin/nvcc").exists():
        pytest.skip("nvcc not available")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = HERE / "fftfilt_emul.cu"
    deps = [src] + list((HERE.parent.parent / "rustradio_b200" / "csrc").glob("fft*"))
    if not SO.exists() or SO.stat().st_mtime < max(d.stat().st_mtime for d in deps):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
                        "-o", str(SO), str(src)], check=True)
    L = C.CDLL(str(SO))
    for fn in (L.emul_fftfilt, L.emul_fftfilt16):
        fn.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p,
                       C.c_longlong, C.c_longlong, C.c_longlong]

    def run(taps, x, hist=None, deci=1, skip=0, variant=32):
        taps = np.ascontiguousarray(taps, np.complex64)
        x = np.ascontiguousarray(x, np.complex64)
        n = len(x)
        n_out = n if (deci == 1 and skip == 0) else ((n - skip + deci - 1) // deci if n > skip else 0)
        out = np.zeros(n_out, np.complex64)
        (L.emul_fftfilt if variant == 32 else L.emul_fftfilt16)(taps.ctypes.data, len(taps), x.ctypes.data, n,
                       hist.ctypes.data if hist is not None else None, out.ctypes.data, deci, skip, n_out)
        return out
    return run


@pytest.mark.parametrize("ntaps,n", [(1, 3000), (2, 20_000), (193, 8000), (4097, 40_000), (64, 16384 * 2 + 5), (12289, 20_000), (16385, 40_000), (20_000, 30_000)])
@pytest.mark.parametrize("variant", [32, 16])
def test_emulated_kernel_matches_f64_convolution(emul, ntaps, n, variant):
    taps = (O.low_pass_n(1.0, 0.05, ntaps).astype(np.complex64) * (1 + 0.3j)) if ntaps > 2 else np.array([0.5 - 0.25j, 0.3 + 1j][:ntaps], np.complex64)
    x = O.synth_c32(5, 0, n)
    assert O.rel_rms(emul(taps, x, variant=variant), O.conv_full_f64_fft(x, taps, n)) <= 1e-5


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("ntaps,n", [(1, 3000), (193, 8000), (4097, 40_000), (64, 16384 * 2 + 5), (16385, 40_000), (4098, 70_000)])
def test_emulated_kernel_twiddles_in_phase_c_and_staged_input(emul, monkeypatch, ntaps, n, mode):
    """fftfilt_core.cuh TWC path (W_512 twiddles from powers inside phase C), stage_input /
    phase_a_staged, linear TMA staging (mode 3) and the PACKED kernel of fftfilt_pk.cuh (mode 4: FFMA2
    lanes, pair-word exchange layouts): same index math and values as the table-twiddle kernel."""
    monkeypatch.setenv("RRC_EMUL_FFTFILT_MODE", str(mode))
    taps = (O.low_pass_n(1.0, 0.05, ntaps).astype(np.complex64) * (1 + 0.3j)) if ntaps > 2 else np.array([0.5 - 0.25j], np.complex64)
    x = O.synth_c32(7, 0, n)
    assert O.rel_rms(emul(taps, x, variant=32), O.conv_full_f64_fft(x, taps, n)) <= 1e-5


@pytest.mark.parametrize("ntaps,n", [(1, 3000), (193, 30_000), (4097, 12288 * 3 + 5), (4097, 12288 * 2), (64, 16384 * 4 + 5), (16385, 70_000)])
def test_emulated_real_stream_mode(emul, ntaps, n):
    """FftFilterFloat mode: two consecutive real blocks per complex transform (odd and even block
    counts, partial last block, tap partitions), with carried f32 history."""
    import ctypes as C
    from pathlib import Path
    L = C.CDLL(str(Path(__file__).parent / "emul" / "_fftfilt_emul.so"))
    L.emul_fftfilt_real.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
    taps = O.low_pass_n(1.0, 0.05, ntaps).astype(np.float32) if ntaps > 1 else np.array([0.75], np.float32)
    x = O.synth_f32(8, 0, n)
    truth = O.conv_full_f64_fft(x.astype(np.complex64), taps.astype(np.complex64), n).real
    out = np.zeros(n, np.float32)
    L.emul_fftfilt_real(taps.ctypes.data, ntaps, x.ctypes.data, n, None, out.ctypes.data)
    assert O.rel_rms(out, truth) <= 1e-5
    cut = n // 3 + 1
    o1, o2 = np.zeros(cut, np.float32), np.zeros(n - cut, np.float32)
    L.emul_fftfilt_real(taps.ctypes.data, ntaps, x.ctypes.data, cut, None, o1.ctypes.data)
    hist = np.zeros(max(ntaps - 1, 1), np.float32)
    if ntaps > 1:
        src = np.concatenate([np.zeros(ntaps - 1, np.float32), x[:cut]])
        hist = np.ascontiguousarray(src[len(src) - (ntaps - 1):])
    x2 = np.ascontiguousarray(x[cut:])
    L.emul_fftfilt_real(taps.ctypes.data, ntaps, x2.ctypes.data, n - cut, hist.ctypes.data if ntaps > 1 else None, o2.ctypes.data)
    assert O.rel_rms(np.concatenate([o1, o2]), truth) <= 1e-5


@pytest.mark.parametrize("variant", [32, 16])
def test_emulated_kernel_history_and_decimation(emul, variant):
    taps = O.low_pass_n(1.0, 0.1, 301).astype(np.complex64)
    x = O.synth_c32(6, 0, 40_000)
    truth = O.conv_full_f64_fft(x, taps, len(x))
    y = np.concatenate([emul(taps, x[:12345], variant=variant), emul(taps, x[12345:], hist=np.ascontiguousarray(x[12345 - 300:12345]), variant=variant)])
    assert O.rel_rms(y, truth) <= 1e-5
    t2 = (O.low_pass_n(1.0, 0.02, 16385).astype(np.complex64) * (1 - 0.2j))
    truth2 = O.conv_full_f64_fft(x, t2, len(x))
    assert O.rel_rms(emul(t2, x, deci=8, skip=0, variant=variant), truth2[::8]) <= 1e-5
    for deci, skip in ((8, 3), (3, 0), (1, 7), (1000, 999), (700, 40_001)):
        yd = emul(taps, x, deci=deci, skip=skip, variant=variant)
        want = truth[skip::deci]
        assert len(yd) == len(want)
        if len(want):
            assert O.rel_rms(yd, want) <= 1e-5


@pytest.fixture(scope="module")
def emul_fold(emul):
    L = C.CDLL(str(SO))
    L.emul_fftfilt_fold.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p,
                                    C.c_longlong, C.c_longlong]

    def run(nc, taps, x, hist=None, skip=0):
        taps = np.ascontiguousarray(taps, np.complex64)
        x = np.ascontiguousarray(x, np.complex64)
        n = len(x)
        n_out = (n - skip + 7) // 8 if n > skip else 0
        out = np.full(n_out, np.nan + 0j, np.complex64)
        L.emul_fftfilt_fold(nc, taps.ctypes.data, len(taps), x.ctypes.data, n,
                            hist.ctypes.data if hist is not None else None, out.ctypes.data, skip, n_out)
        return out
    return run


@pytest.mark.parametrize("nc,ntaps,n,skip", [(1, 301, 40_000, 0), (1, 4097, 50_000, 3), (1, 1, 20_000, 7), (1, 12289, 30_000, 13),
                                             (4, 16385, 150_000, 0), (4, 16385, 70_000, 5), (4, 12290, 120_000, 9),
                                             (4, 40_001, 100_000, 2)])
def test_emulated_fold_kernel_matches_f64_convolution(emul_fold, nc, ntaps, n, skip):
    """Decimate-by-8 fold kernel (pruned inverse; nc = 4: 65536-point cluster geometry)."""
    taps = (O.low_pass_n(1.0, 0.02, ntaps).astype(np.complex64) * (1 - 0.2j)) if ntaps > 2 else np.array([0.5 - 0.25j], np.complex64)
    x = O.synth_c32(7, 0, n)
    want = O.conv_full_f64_fft(x, taps, n)[skip::8]
    got = emul_fold(nc, taps, x, skip=skip)
    assert len(got) == len(want) and not np.isnan(got).any()
    assert O.rel_rms(got, want) <= 1e-5


def test_emulated_fold_kernel_streaming_history(emul_fold):
    taps = O.low_pass_n(1.0, 0.02, 16385).astype(np.complex64)
    x = O.synth_c32(8, 0, 120_000)
    truth = O.conv_full_f64_fft(x, taps, len(x))
    cut = 70_001
    a = emul_fold(4, taps, x[:cut], skip=0)
    skip2 = (-cut) % 8                      # RationalResampler(1, 8) phase carried across the call boundary
    b = emul_fold(4, taps, x[cut:], hist=np.ascontiguousarray(x[cut - 16384:cut]), skip=skip2)
    assert O.rel_rms(np.concatenate([a, b]), truth[::8]) <= 1e-5


@pytest.fixture(scope="module")
def emul_poly(emul):
    L = C.CDLL(str(SO))
    L.emul_fftfilt_poly.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p,
                                    C.c_longlong, C.c_longlong, C.c_longlong]

    def run(taps, x, deci, hist=None, skip=0):
        taps = np.ascontiguousarray(taps, np.complex64)
        x = np.ascontiguousarray(x, np.complex64)
        n = len(x)
        n_out = (n - skip + deci - 1) // deci if n > skip else 0
        out = np.full(n_out, np.nan + 0j, np.complex64)
        L.emul_fftfilt_poly(taps.ctypes.data, len(taps), x.ctypes.data, n,
                            hist.ctypes.data if hist is not None else None, out.ctypes.data, deci, skip, n_out)
        return out
    return run


@pytest.mark.parametrize("ntaps,n,deci,skip", [(16385, 300_000, 8, 0), (16385, 70_000, 8, 5), (301, 40_000, 8, 0), (4097, 50_000, 8, 3),
                                               (5, 20_000, 8, 7), (40_001, 200_000, 8, 2), (4097, 60_000, 3, 1), (64, 50_000, 2, 0),
                                               (16385, 90_000, 16, 9), (12289, 60_000, 5, 13)])
def test_emulated_poly_kernel_matches_f64_convolution(emul_poly, ntaps, n, deci, skip):
    """Polyphase decimating kernel (fftfilt_poly_core.cuh): deci forward transforms per block, the sum over the branches in
    the per-thread accumulator, one inverse transform."""
    taps = O.low_pass_n(1.0, 0.02, ntaps).astype(np.complex64) * (1 - 0.2j)
    x = O.synth_c32(7, 0, n)
    want = O.conv_full_f64_fft(x, taps, n)[skip::deci]
    got = emul_poly(taps, x, deci, skip=skip)
    assert len(got) == len(want) and not np.isnan(got).any()
    assert O.rel_rms(got, want) <= 1e-5


def test_emulated_poly_kernel_streaming_history(emul_poly):
    taps = O.low_pass_n(1.0, 0.02, 16385).astype(np.complex64)
    x = O.synth_c32(8, 0, 150_000)
    truth = O.conv_full_f64_fft(x, taps, len(x))
    cut = 70_001
    a = emul_poly(taps, x[:cut], 8, skip=0)
    skip2 = (-cut) % 8                      # RationalResampler(1, 8) phase carried across the call boundary
    b = emul_poly(taps, x[cut:], 8, hist=np.ascontiguousarray(x[cut - 16384:cut]), skip=skip2)
    assert O.rel_rms(np.concatenate([a, b]), truth[::8]) <= 1e-5
