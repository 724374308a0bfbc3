"""ctypes/numpy front end of the CPU oracle (oracle/rr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(rustradio_b200/) must never import this module.

Every function is a thin wrapper; the algorithm and its reference citations
live in rr_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_BUILD = _HERE / "_build"

WINDOW_HAMMING, WINDOW_BLACKMAN, WINDOW_BLACKMAN_HARRIS, WINDOW_HAMMING_PARM = 0, 1, 2, 3


def build(force: bool = False) -> None:
    """Compile the oracle with gcc (oracle/Makefile)."""
    so = _BUILD / "librr_oracle.so"
    src = _HERE / "rr_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)


def _load(name: str) -> C.CDLL:
    path = _BUILD / name
    if not path.exists():
        build()
    lib = C.CDLL(str(path))
    i64, f32, f64, vp = C.c_int64, C.c_float, C.c_double, C.c_void_p
    sig = {
        "orc_fir_out_count": (i64, [i64, i64, i64]),
        "orc_fir_c32": (None, [vp, vp, i64, i64, vp, i64]),
        "orc_fir_f32": (None, [vp, vp, i64, i64, vp, i64]),
        "orc_fir_c32_f64": (None, [vp, vp, i64, i64, vp, i64]),
        "orc_fir_f32_f64": (None, [vp, vp, i64, i64, vp, i64]),
        "orc_fir_new_translator": (C.c_int, [vp, i64, f32, f32, i64, vp, vp]),
        "orc_fir_translate_output": (None, [vp, i64, vp, vp]),
        "orc_make_window": (C.c_int, [C.c_int, f32, i64, vp]),
        "orc_compute_ntaps": (i64, [f32, f32, C.c_int]),
        "orc_low_pass_n": (None, [f32, f32, C.c_int, f32, i64, vp]),
        "orc_calc_fft_size": (i64, [i64]),
        "orc_fftfilt_out_count": (i64, [i64, i64]),
        "orc_fft_c32": (None, [vp, i64, C.c_int]),
        "orc_fftfilt_new": (vp, [vp, i64]),
        "orc_fftfilt_free": (None, [vp]),
        "orc_fftfilt_nsamples": (i64, [vp]),
        "orc_fftfilt_fft_size": (i64, [vp]),
        "orc_fftfilt_run": (None, [vp, vp, i64, vp]),
        "orc_conv_full_c32_f64": (None, [vp, i64, vp, i64, vp]),
        "orc_resampler_new": (vp, [i64, i64, i64]),
        "orc_resampler_free": (None, [vp]),
        "orc_resampler_has_pending": (C.c_int, [vp]),
        "orc_resampler_counter": (i64, [vp]),
        "orc_resampler_set_counter": (None, [vp, i64]),
        "orc_resampler_work": (C.c_int, [vp, vp, i64, vp, i64, C.POINTER(i64), C.POINTER(i64)]),
        "orc_resample_out_count": (i64, [i64, i64, i64]),
        "orc_quad_demod": (None, [vp, i64, f32, vp]),
        "orc_quad_demod_f64": (None, [vp, i64, f64, vp]),
        "orc_rtlsdr_decode": (None, [vp, i64, vp]),
        "orc_rtlsdr_encode": (None, [vp, i64, vp]),
        "orc_hilbert_taps": (C.c_int, [vp, i64, vp]),
        "orc_hilbert_work": (None, [vp, vp, i64, vp, i64, vp, vp]),
        "orc_multiply_const_f32": (None, [vp, i64, f32, vp]),
        "orc_multiply_const_c32": (None, [vp, i64, f32, f32, vp]),
        "orc_add_const_f32": (None, [vp, i64, f32, vp]),
        "orc_add_const_c32": (None, [vp, i64, f32, f32, vp]),
        "orc_complex_to_mag2": (None, [vp, i64, vp]),
        "orc_iq_balance_alpha_from_tau": (f32, [C.c_uint32, f64]),
        "orc_iq_balance": (None, [vp, i64, f32, vp, vp]),
        "orc_iq_balance_f64": (None, [vp, i64, f32, vp, vp]),
        "orc_signal_source_complex": (None, [f32, f32, f32, C.POINTER(f64), vp, i64]),
        "orc_synth_f32": (None, [C.c_uint64, C.c_uint64, vp, i64]),
    }
    for fn, (res, args) in sig.items():
        f = getattr(lib, fn)
        f.restype = res
        f.argtypes = args
    return lib


_libs: dict[str, C.CDLL] = {}


def lib(fast: bool = False) -> C.CDLL:
    """faithful (-O2) build by default; fast=True is the -O3 AVX2 baseline build."""
    name = "librr_oracle_fast.so" if fast else "librr_oracle.so"
    if name not in _libs:
        _libs[name] = _load(name)
    return _libs[name]


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def _c64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.complex64)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


# ---------------------------------------------------------------- FIR -----
def fir_out_count(n: int, ntaps: int, deci: int = 1) -> int:
    return int(lib().orc_fir_out_count(n, ntaps, deci))


def fir(x, taps, deci: int = 1, *, f64: bool = False, fast: bool = False) -> np.ndarray:
    """Whole-stream FirFilter output (valid windows only)."""
    cplx = np.iscomplexobj(x) or np.iscomplexobj(taps)
    L = lib(fast)
    if cplx:
        x, taps = _c64(x), _c64(taps)
        n_out = fir_out_count(len(x), len(taps), deci)
        out = np.empty(n_out, dtype=np.complex128 if f64 else np.complex64)
        if n_out:
            (L.orc_fir_c32_f64 if f64 else L.orc_fir_c32)(_p(x), _p(taps), len(taps), deci, _p(out), n_out)
    else:
        x, taps = _f32(x), _f32(taps)
        n_out = fir_out_count(len(x), len(taps), deci)
        out = np.empty(n_out, dtype=np.float64 if f64 else np.float32)
        if n_out:
            (L.orc_fir_f32_f64 if f64 else L.orc_fir_f32)(_p(x), _p(taps), len(taps), deci, _p(out), n_out)
    return out


def fir_new_translator(taps, samp_rate: float, freq: float, deci: int):
    """Returns (rotated_taps, phase0, step) or (taps, None, None) if freq == 0."""
    t = _c64(taps).copy()
    ph = np.zeros(1, np.complex64)
    st = np.zeros(1, np.complex64)
    ok = lib().orc_fir_new_translator(_p(t), len(t), samp_rate, freq, deci, _p(ph), _p(st))
    if not ok:
        return t, None, None
    return t, ph[0], st[0]


def fir_translate_output(out, phase, step):
    """In-place rotate; returns the carried phase."""
    o = out  # must be contiguous complex64
    assert o.dtype == np.complex64 and o.flags.c_contiguous
    ph = np.array([phase], np.complex64)
    st = np.array([step], np.complex64)
    lib().orc_fir_translate_output(_p(o), len(o), _p(ph), _p(st))
    return ph[0]


# --------------------------------------------------------- tap design -----
def make_window(window_type: int, ntaps: int, parm: float = 0.0) -> np.ndarray:
    w = np.empty(ntaps, np.float32)
    lib().orc_make_window(window_type, parm, ntaps, _p(w))
    return w


def compute_ntaps(samp_rate: float, twidth: float, window_type: int = WINDOW_HAMMING) -> int:
    return int(lib().orc_compute_ntaps(samp_rate, twidth, window_type))


def low_pass_n(samp_rate: float, cutoff: float, ntaps: int, window_type: int = WINDOW_HAMMING,
               parm: float = 0.0) -> np.ndarray:
    t = np.empty(ntaps, np.float32)
    lib().orc_low_pass_n(samp_rate, cutoff, window_type, parm, ntaps, _p(t))
    return t


def low_pass(samp_rate: float, cutoff: float, twidth: float, window_type: int = WINDOW_HAMMING,
             parm: float = 0.0) -> np.ndarray:
    """rustradio::fir::low_pass (src/fir.rs:617-656)."""
    return low_pass_n(samp_rate, cutoff, compute_ntaps(samp_rate, twidth, window_type), window_type, parm)


def low_pass_complex(samp_rate, cutoff, twidth, window_type=WINDOW_HAMMING, parm=0.0) -> np.ndarray:
    return low_pass(samp_rate, cutoff, twidth, window_type, parm).astype(np.complex64)


# ---------------------------------------------------------- FFT filter ----
def calc_fft_size(ntaps: int) -> int:
    return int(lib().orc_calc_fft_size(ntaps))


def fftfilt_out_count(n: int, ntaps: int) -> int:
    return int(lib().orc_fftfilt_out_count(n, ntaps))


def fft(x, inverse: bool = False) -> np.ndarray:
    b = _c64(x).copy()
    lib().orc_fft_c32(_p(b), len(b), int(inverse))
    return b


class FftFilt:
    """Stateful restatement of FftFilter's overlap-add engine (whole blocks)."""

    def __init__(self, taps, fast: bool = False):
        self._L = lib(fast)
        t = _c64(taps)
        self._h = self._L.orc_fftfilt_new(_p(t), len(t))
        self.ntaps = len(t)
        self.nsamples = int(self._L.orc_fftfilt_nsamples(self._h))
        self.fft_size = int(self._L.orc_fftfilt_fft_size(self._h))

    def run(self, x) -> np.ndarray:
        x = _c64(x)
        nb = len(x) // self.nsamples
        out = np.empty(nb * self.nsamples, np.complex64)
        if nb:
            self._L.orc_fftfilt_run(self._h, _p(x), nb, _p(out))
        return out

    def __del__(self):
        try:
            self._L.orc_fftfilt_free(self._h)
        except Exception:
            pass


def fftfilt(x, taps, fast: bool = False) -> np.ndarray:
    """Whole-stream FftFilter output: floor(N/S)*S samples of the full convolution."""
    return FftFilt(taps, fast).run(x)


def conv_full_f64(x, taps, n_out: int) -> np.ndarray:
    """f64 truth y[n] = sum_k h[k] x[n-k], x[n<0]=0, first n_out samples (direct)."""
    x, t = _c64(x), _c64(taps)
    out = np.empty(n_out, np.complex128)
    lib().orc_conv_full_c32_f64(_p(x), n_out, _p(t), len(t), _p(out))
    return out


def conv_full_f64_fft(x, taps, n_out: int) -> np.ndarray:
    """Same truth via numpy's f64 FFT (for sizes where direct is too slow)."""
    x = np.asarray(x).astype(np.complex128)[:n_out]
    t = np.asarray(taps).astype(np.complex128)
    n = 1
    while n < len(x) + len(t):
        n <<= 1
    y = np.fft.ifft(np.fft.fft(x, n) * np.fft.fft(t, n))
    return y[:n_out]


# ----------------------------------------------------------- resampler ----
_DT_BY_SIZE = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}


class Resampler:
    """Stateful restatement of RationalResampler::work()."""
    WAIT_DST, WAIT_SRC = 0, 1

    def __init__(self, elem_size: int, interp: int, deci: int):
        self._h = lib().orc_resampler_new(elem_size, interp, deci)
        if not self._h:
            raise ValueError("RationalResampler created using interp/deci 0")
        self.elem = elem_size

    def work(self, inp: np.ndarray, out_cap: int):
        """Returns (ret, consumed, out_array)."""
        inp = np.ascontiguousarray(inp)
        assert inp.dtype.itemsize == self.elem
        out = np.empty(out_cap, inp.dtype)
        c, p = C.c_int64(0), C.c_int64(0)
        ret = lib().orc_resampler_work(self._h, _p(inp), len(inp), _p(out), out_cap, C.byref(c), C.byref(p))
        return ret, int(c.value), out[: p.value].copy()

    def set_counter(self, counter: int):
        """Preset the reference's `counter` field (mid-stream start of a time-segment shard)."""
        lib().orc_resampler_set_counter(self._h, counter)

    @property
    def has_pending(self) -> bool:
        return bool(lib().orc_resampler_has_pending(self._h))

    @property
    def counter(self) -> int:
        return int(lib().orc_resampler_counter(self._h))

    def __del__(self):
        try:
            lib().orc_resampler_free(self._h)
        except Exception:
            pass


def resample_out_count(n: int, interp: int, deci: int) -> int:
    return int(lib().orc_resample_out_count(n, interp, deci))


def resample(x: np.ndarray, interp: int, deci: int) -> np.ndarray:
    """Whole-stream output (unbounded output window)."""
    x = np.ascontiguousarray(x)
    r = Resampler(x.dtype.itemsize, interp, deci)
    cap = resample_out_count(len(x), interp, deci) + 1
    _, consumed, out = r.work(x, cap)
    assert consumed == len(x)
    return out


# --------------------------------------------------------------- demod ----
def quad_demod(x, gain: float = 1.0, *, f64: bool = False, fast: bool = False) -> np.ndarray:
    x = _c64(x)
    n = max(len(x) - 1, 0)
    out = np.empty(n, np.float64 if f64 else np.float32)
    if n:
        if f64:
            lib(fast).orc_quad_demod_f64(_p(x), len(x), gain, _p(out))
        else:
            lib(fast).orc_quad_demod(_p(x), len(x), gain, _p(out))
    return out


# -------------------------------------------------------- rtlsdr decode ----
def rtlsdr_decode(raw) -> np.ndarray:
    """RtlSdrDecode (src/rtlsdr_decode.rs:35-43): u8 I/Q pairs -> c32, (b - 127) * 0.008."""
    raw = np.ascontiguousarray(raw, np.uint8)
    out = np.empty(len(raw) // 2, np.complex64)
    if len(out):
        lib().orc_rtlsdr_decode(_p(raw), len(raw), _p(out))
    return out


def rtlsdr_encode(x) -> np.ndarray:
    """RtlSdrEncode (src/rtlsdr_encode.rs:22-26): c32 -> u8 I/Q pairs, round((s / 0.008) + 127) clamped to 0..255."""
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty(2 * len(x), np.uint8)
    if len(x):
        lib().orc_rtlsdr_encode(_p(x), len(x), _p(out))
    return out


def synth_u8(seed: int, first_index: int, n: int) -> np.ndarray:
    """Deterministic u8 I/Q bytes: the top byte of the same splitmix64 stream synth_f32 uses."""
    f = synth_f32(seed, first_index, n)
    return np.clip(np.floor((f.astype(np.float64) + 1.0) * 128.0), 0, 255).astype(np.uint8)


# -------------------------------------------------------------- Hilbert ----
def hilbert_taps(window) -> np.ndarray:
    """fir::hilbert (src/fir.rs:660-680)."""
    w = _f32(window)
    t = np.empty(len(w), np.float32)
    if lib().orc_hilbert_taps(_p(w), len(w), _p(t)) != 0:
        raise ValueError("hilbert() needs a window of at least 2 taps")
    return t


class Hilbert:
    """Hilbert::work compute with its carried history (src/hilbert.rs:52,86-125)."""

    def __init__(self, ntaps: int, window_type: int = WINDOW_HAMMING, parm: float = 0.0, taps=None):
        assert ntaps > 1 and ntaps & 1 == 1, "hilbert filter len must be odd and greater than 1"   # :44-47
        self.ntaps = ntaps
        # taps=: test hook for arbitrary (non half-band) taps; the block always uses fir::hilbert
        self.taps = hilbert_taps(make_window(window_type, ntaps, parm)) if taps is None else _f32(taps)
        assert len(self.taps) == ntaps
        self.history = np.zeros(ntaps, np.float32)

    def work(self, x, *, f64: bool = False) -> np.ndarray:
        x = _f32(x)
        out = np.empty(len(x), np.complex128 if f64 else np.complex64)
        if len(x):
            lib().orc_hilbert_work(_p(self.history), _p(self.taps), self.ntaps, _p(x), len(x),
                                   None if f64 else _p(out), _p(out) if f64 else None)
        return out


# ------------------------------------------------ sample-wise neighbours ----
def multiply_const(x, val) -> np.ndarray:
    if np.iscomplexobj(x):
        x = _c64(x); out = np.empty_like(x); v = complex(val)
        lib().orc_multiply_const_c32(_p(x), len(x), v.real, v.imag, _p(out))
    else:
        x = _f32(x); out = np.empty_like(x)
        lib().orc_multiply_const_f32(_p(x), len(x), float(val), _p(out))
    return out


def add_const(x, val) -> np.ndarray:
    if np.iscomplexobj(x):
        x = _c64(x); out = np.empty_like(x); v = complex(val)
        lib().orc_add_const_c32(_p(x), len(x), v.real, v.imag, _p(out))
    else:
        x = _f32(x); out = np.empty_like(x)
        lib().orc_add_const_f32(_p(x), len(x), float(val), _p(out))
    return out


def complex_to_mag2(x) -> np.ndarray:
    x = _c64(x)
    out = np.empty(len(x), np.float32)
    lib().orc_complex_to_mag2(_p(x), len(x), _p(out))
    return out


def iq_balance_alpha_from_tau(sample_rate: int, tau_seconds: float = 0.2) -> float:
    return float(lib().orc_iq_balance_alpha_from_tau(sample_rate, tau_seconds))


class IqBalance:
    """IqBalance::process_sync with its carried mean (src/iq_balance.rs:62-80)."""

    def __init__(self, alpha: float):
        self.alpha = float(np.float32(min(max(alpha, 0.0), 1.0)))
        self.mean = np.zeros(1, np.complex64)
        self.mean64 = np.zeros(1, np.complex128)

    def work(self, x, *, f64: bool = False) -> np.ndarray:
        x = _c64(x)
        if f64:
            out = np.empty(len(x), np.complex128)
            lib().orc_iq_balance_f64(_p(x), len(x), self.alpha, _p(self.mean64), _p(out))
        else:
            out = np.empty(len(x), np.complex64)
            lib().orc_iq_balance(_p(x), len(x), self.alpha, _p(self.mean), _p(out))
        return out


# ------------------------------------------------------------ fixtures ----
def signal_source_complex(samp_rate: float, freq: float, amplitude: float, n: int, current: float = 0.0):
    out = np.empty(n, np.complex64)
    cur = C.c_double(current)
    lib().orc_signal_source_complex(samp_rate, freq, amplitude, C.byref(cur), _p(out), n)
    return out, cur.value


def synth_f32(seed: int, first_index: int, n: int) -> np.ndarray:
    out = np.empty(n, np.float32)
    lib().orc_synth_f32(seed, first_index, _p(out), n)
    return out


def synth_c32(seed: int, first_sample: int, n: int) -> np.ndarray:
    """Complex white noise, re/im ~ U(-1,1): float index 2*s is re, 2*s+1 is im."""
    return synth_f32(seed, 2 * first_sample, 2 * n).view(np.complex64)


# ------------------------------------------------------------- metrics ----
def rel_rms(y, ref) -> float:
    y = np.asarray(y)
    ref = np.asarray(ref)
    d = np.linalg.norm(y.astype(np.complex128) - ref.astype(np.complex128))
    n = np.linalg.norm(ref.astype(np.complex128))
    return float(d / n) if n > 0 else float(d)


def max_angle_err(a, ref) -> float:
    """max |a-ref| modulo 2*pi (demod bar is in radians)."""
    d = np.asarray(a, np.float64) - np.asarray(ref, np.float64)
    d = (d + np.pi) % (2 * np.pi) - np.pi
    return float(np.max(np.abs(d))) if d.size else 0.0
