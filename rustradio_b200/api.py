"""ctypes binding of include/rustradio_cuda.h (see package docstring)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "librustradio_cuda.so"

RRC_OK = 0
RRC_FIR_NO_REAL_TAP_FASTPATH = 1
RRC_FIR_FORCE_GENERIC = 2
RRC_FIR_NO_TENSOR = 4
EPI_NONE, EPI_MULTIPLY_CONST, EPI_ADD_CONST, EPI_MAG2 = 0, 1, 2, 3


class RrcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rrc error {code}: {msg}")
        self.code = code


def library_path() -> Path:
    return _SO


def build_library(verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", str(_HERE / "csrc"), "-j8"] + ([] if verbose else ["-s"]), check=True)
    return _SO


_lib = None

# name -> argtypes (restype is always int except rrc_last_error)
_sz, _vp, _i, _u64, _f = C.c_size_t, C.c_void_p, C.c_int, C.c_uint64, C.c_float
_P = C.POINTER
_SIGS = {
    "rrc_abi_version": [],
    "rrc_device_count": [_P(_i)],
    "rrc_device_name": [_i, C.c_char_p, _sz],
    "rrc_device_sm_count": [_i, _P(_i)],
    "rrc_malloc_device": [_i, _sz, _P(_vp)],
    "rrc_free_device": [_i, _vp],
    "rrc_malloc_pinned": [_sz, _P(_vp)],
    "rrc_free_pinned": [_vp],
    "rrc_malloc_pinned_near": [_i, _sz, _P(_vp)],
    "rrc_device_numa_node": [_i, _P(_i)],
    "rrc_peer_enable": [_i, _i],
    "rrc_memcpy_peer": [_i, _vp, _i, _vp, _sz, _vp],
    "rrc_ipc_export": [_i, _vp, _vp],
    "rrc_ipc_open": [_i, _vp, _P(_vp)],
    "rrc_ipc_close": [_i, _vp],
    "rrc_host_register": [_vp, _sz],
    "rrc_host_unregister": [_vp],
    "rrc_memset_device": [_i, _vp, _i, _sz, _vp],
    "rrc_memcpy_h2d": [_i, _vp, _vp, _sz, _vp],
    "rrc_memcpy_d2h": [_i, _vp, _vp, _sz, _vp],
    "rrc_memcpy_d2d": [_i, _vp, _vp, _sz, _vp],
    "rrc_stream_create": [_i, _P(_vp)],
    "rrc_stream_destroy": [_i, _vp],
    "rrc_stream_sync": [_i, _vp],
    "rrc_device_sync": [_i],
    "rrc_event_create": [_i, _P(_vp)],
    "rrc_event_destroy": [_i, _vp],
    "rrc_event_record": [_i, _vp, _vp],
    "rrc_event_wait": [_i, _vp, _vp],
    "rrc_event_sync": [_i, _vp],
    "rrc_event_elapsed_ms": [_i, _vp, _vp, _P(_f)],
    "rrc_synth_f32": [_i, _u64, _u64, _vp, _sz, _vp],
    "rrc_launch_count": [_P(_u64)],
    "rrc_fir_c32_create": [_i, _vp, _sz, _sz, C.c_uint, _P(_vp)],
    "rrc_fir_f32_create": [_i, _vp, _sz, _sz, C.c_uint, _P(_vp)],
    "rrc_fir_set_translate": [_vp, _f, _f],
    "rrc_fir_destroy": [_vp],
    "rrc_fir_ntaps": [_vp, _P(_sz)],
    "rrc_fir_deci": [_vp, _P(_sz)],
    "rrc_fir_uses_real_taps": [_vp, _P(_i)],
    "rrc_fir_uses_tensor_cores": [_vp, _P(_i)],
    "rrc_fir_reset": [_vp],
    "rrc_fir_kernel_name": [_vp, C.c_char_p, _sz],
    "rrc_fir_set_epilogue": [_vp, _i, _f, _f],
    "rrc_fftfilt_set_epilogue": [_vp, _i, _f, _f],
    "rrc_fir_plan": [_sz, _sz, _sz, _sz, _P(_sz), _P(_sz), _P(_sz), _P(_sz), _P(_i)],
    "rrc_fir_run": [_vp, _vp, _sz, _vp, _sz, _vp],
    "rrc_fir_run_batch": [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _sz, _vp],
    "rrc_fir_c32_demod_run_batch": [_vp, _vp, _sz, _sz, _f, _vp, _sz, _sz, _sz, _vp],
    "rrc_fir_run_host": [_vp, _vp, _sz, _vp, _P(_sz)],
    "rrc_fir_c32_demod_run_host_batch": [_vp, _vp, _sz, _sz, _f, _vp, _sz, _P(_sz)],
    "rrc_fftfilt_c32_create": [_i, _vp, _sz, _P(_vp)],
    "rrc_fftfilt_f32_create": [_i, _vp, _sz, _P(_vp)],
    "rrc_fftfilt_destroy": [_vp],
    "rrc_fftfilt_reset": [_vp, _vp],
    "rrc_fftfilt_set_history": [_vp, _vp, _sz, _vp],
    "rrc_fftfilt_set_history_ptr": [_vp, _vp, _sz],
    "rrc_fftfilt_ref_fft_size": [_sz, _P(_sz), _P(_sz)],
    "rrc_fftfilt_geometry": [_vp, _P(_sz), _P(_sz)],
    "rrc_fftfilt_plan": [_sz, _sz, _sz, _sz, _P(_sz), _P(_sz), _P(_sz), _P(_sz), _P(_i)],
    "rrc_fftfilt_run": [_vp, _vp, _sz, _vp, _vp],
    "rrc_fftfilt_decim_run": [_vp, _vp, _sz, _sz, _sz, _vp, _P(_sz), _vp],
    "rrc_fftfilt_run_host": [_vp, _vp, _sz, _vp, _P(_sz)],
    "rrc_fftfilt_decim_run_host": [_vp, _vp, _sz, _sz, _vp, _P(_sz)],
    "rrc_resampler_create": [_i, _sz, _sz, _sz, _P(_vp)],
    "rrc_resampler_destroy": [_vp],
    "rrc_resampler_reset": [_vp],
    "rrc_resampler_set_state": [_vp, C.c_int64, _vp],
    "rrc_resampler_state": [_vp, _P(C.c_int64), _P(C.c_int64), _P(C.c_int64), _P(_i)],
    "rrc_resampler_run": [_vp, _vp, _sz, _vp, _sz, _P(_sz), _P(_sz), _P(_i), _vp],
    "rrc_resampler_run_host": [_vp, _vp, _sz, _vp, _sz, _P(_sz), _P(_sz)],
    "rrc_quad_demod_run": [_i, _vp, _sz, _f, _vp, _vp],
    "rrc_quad_demod_run_batch": [_i, _vp, _sz, _sz, _f, _vp, _sz, _sz, _vp],
    "rrc_quad_demod_run_host": [_i, _vp, _sz, _f, _vp],
    "rrc_fft_c32_create": [_i, _sz, _P(_vp)],
    "rrc_fft_destroy": [_vp],
    "rrc_fft_size": [_vp, _P(_sz)],
    "rrc_fft_run": [_vp, _vp, _sz, _vp, _vp],
    "rrc_fftstream_plan": [_sz, _sz, _sz, _P(_sz), _P(_sz), _P(_i)],
    "rrc_fft_run_host": [_vp, _vp, _sz, _vp, _P(_sz)],
    "rrc_rtlsdr_decode_plan": [_sz, _sz, _P(_sz), _P(_sz), _P(_sz), _P(_i)],
    "rrc_rtlsdr_decode_run": [_i, _vp, _sz, _vp, _vp],
    "rrc_rtlsdr_decode_run_host": [_i, _vp, _sz, _vp, _P(_sz)],
    "rrc_rtlsdr_encode_plan": [_sz, _sz, _P(_sz), _P(_sz), _P(_sz), _P(_i)],
    "rrc_rtlsdr_encode_run": [_i, _vp, _sz, _vp, _vp],
    "rrc_rtlsdr_encode_run_host": [_i, _vp, _sz, _vp, _P(_sz)],
    "rrc_fir_set_input_u8iq": [_vp, _i],
    "rrc_fftfilt_set_input_u8iq": [_vp, _i],
    "rrc_make_window": [_i, _f, _sz, _vp],
    "rrc_hilbert_taps": [_vp, _sz, _vp],
    "rrc_hilbert_create": [_i, _vp, _sz, _P(_vp)],
    "rrc_hilbert_destroy": [_vp],
    "rrc_hilbert_reset": [_vp, _vp],
    "rrc_hilbert_run": [_vp, _vp, _sz, _vp, _vp],
    "rrc_multiply_const_f32_run": [_i, _vp, _sz, _f, _vp, _vp],
    "rrc_multiply_const_c32_run": [_i, _vp, _sz, _f, _f, _vp, _vp],
    "rrc_add_const_f32_run": [_i, _vp, _sz, _f, _vp, _vp],
    "rrc_add_const_c32_run": [_i, _vp, _sz, _f, _f, _vp, _vp],
    "rrc_complex_to_mag2_run": [_i, _vp, _sz, _vp, _vp],
    "rrc_tee_run": [_i, _vp, _sz, _vp, _vp, _vp],
    "rrc_iq_balance_alpha_from_tau": [C.c_uint, C.c_double, _P(_f)],
    "rrc_iq_balance_create": [_i, _f, _P(_vp)],
    "rrc_iq_balance_destroy": [_vp],
    "rrc_iq_balance_reset": [_vp, _vp],
    "rrc_iq_balance_mean": [_vp, _vp, _vp],
    "rrc_iq_balance_run": [_vp, _vp, _sz, _vp, _vp],
}


def exported_symbols() -> list[str]:
    return sorted(_SIGS) + ["rrc_last_error"]


def lib() -> C.CDLL:
    """Load the CUDA library; never falls back to anything else."""
    global _lib
    if _lib is None:
        if not _SO.exists():
            raise RrcError(-2, f"{_SO} is missing: build it with __graft_entry__.build() / make -C rustradio_b200/csrc")
        L = C.CDLL(str(_SO))
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        L.rrc_last_error.restype = C.c_char_p
        L.rrc_last_error.argtypes = []
        _lib = L
    return _lib


def _ck(code: int) -> None:
    if code != RRC_OK:
        raise RrcError(code, lib().rrc_last_error().decode(errors="replace"))


def _ptr(x) -> int:
    """Device/host pointer from DeviceBuffer / PinnedBuffer / numpy / torch tensor / int."""
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if isinstance(x, (DeviceBuffer, PinnedBuffer)):
        return x.ptr
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


def device_count() -> int:
    n = C.c_int(0)
    _ck(lib().rrc_device_count(C.byref(n)))
    return n.value


def launch_count() -> int:
    n = C.c_uint64(0)
    _ck(lib().rrc_launch_count(C.byref(n)))
    return n.value


def stream_sync(device: int = 0, stream: int = 0) -> None:
    _ck(lib().rrc_stream_sync(device, stream))


def device_sync(device: int = 0) -> None:
    _ck(lib().rrc_device_sync(device))


class Event:
    def __init__(self, device: int = 0):
        self.device = device
        h = C.c_void_p()
        _ck(lib().rrc_event_create(device, C.byref(h)))
        self.h = h.value

    def record(self, stream: int = 0):
        _ck(lib().rrc_event_record(self.device, self.h, stream))

    def sync(self):
        _ck(lib().rrc_event_sync(self.device, self.h))

    def elapsed_ms(self, later: "Event") -> float:
        ms = C.c_float(0)
        _ck(lib().rrc_event_elapsed_ms(self.device, self.h, later.h, C.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            lib().rrc_event_destroy(self.device, self.h)
        except Exception:
            pass


def event_wait(event: "Event", stream: int = 0) -> None:
    _ck(lib().rrc_event_wait(event.device, event.h, stream))


def device_numa_node(device: int = 0) -> int:
    n = _i(0)
    _ck(lib().rrc_device_numa_node(device, C.byref(n)))
    return n.value


def peer_enable(device: int, peer: int) -> None:
    _ck(lib().rrc_peer_enable(device, peer))


def ipc_export(buf) -> bytes:
    """64-byte CUDA IPC handle of a DeviceBuffer (cudaMalloc allocation) for another process on this node."""
    h = (C.c_ubyte * 64)()
    _ck(lib().rrc_ipc_export(buf.device, buf.ptr, h))
    return bytes(h)


class IpcMapping:
    """A peer process's device buffer mapped into this process (reads/writes go over NVLink)."""

    def __init__(self, handle: bytes, device: int = 0):
        assert len(handle) == 64
        self.device = device
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        _ck(lib().rrc_ipc_open(device, h, C.byref(p)))
        self.ptr = p.value

    def close(self):
        if self.ptr:
            lib().rrc_ipc_close(self.device, self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceBuffer:
    """Raw device allocation owned by the C library."""

    def __init__(self, nbytes: int, device: int = 0):
        self.device, self.nbytes = device, int(nbytes)
        p = C.c_void_p()
        _ck(lib().rrc_malloc_device(device, self.nbytes, C.byref(p)))
        self.ptr = p.value or 0

    @classmethod
    def from_numpy(cls, a: np.ndarray, device: int = 0) -> "DeviceBuffer":
        a = np.ascontiguousarray(a)
        b = cls(max(a.nbytes, 1), device)
        b.upload(a)
        return b

    def upload(self, a: np.ndarray, offset_bytes: int = 0, stream: int = 0):
        a = np.ascontiguousarray(a)
        assert offset_bytes + a.nbytes <= self.nbytes
        _ck(lib().rrc_memcpy_h2d(self.device, self.ptr + offset_bytes, a.ctypes.data, a.nbytes, stream))
        _ck(lib().rrc_stream_sync(self.device, stream))

    def download(self, dtype, count: int, offset_bytes: int = 0, stream: int = 0) -> np.ndarray:
        out = np.empty(count, dtype)
        assert offset_bytes + out.nbytes <= self.nbytes
        if out.nbytes:
            _ck(lib().rrc_memcpy_d2h(self.device, out.ctypes.data, self.ptr + offset_bytes, out.nbytes, stream))
        _ck(lib().rrc_stream_sync(self.device, stream))
        return out

    def zero(self, stream: int = 0):
        _ck(lib().rrc_memset_device(self.device, self.ptr, 0, self.nbytes, stream))

    def free(self):
        if self.ptr:
            lib().rrc_free_device(self.device, self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy array."""

    def __init__(self, dtype, count: int, near_device: int | None = None):
        """near_device: place the pages on that GPU's NUMA node (rrc_malloc_pinned_near)."""
        self.dtype = np.dtype(dtype)
        self.count = int(count)
        p = C.c_void_p()
        nbytes = max(self.count * self.dtype.itemsize, 1)
        _ck(lib().rrc_malloc_pinned(nbytes, C.byref(p)) if near_device is None else lib().rrc_malloc_pinned_near(near_device, nbytes, C.byref(p)))
        self.ptr = p.value
        buf = (C.c_char * (self.count * self.dtype.itemsize)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=self.count)

    def free(self):
        if self.ptr:
            self.array = None
            lib().rrc_free_pinned(self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def synth_f32(dst, seed: int, first_index: int, n_floats: int, device: int = 0, stream: int = 0) -> None:
    _ck(lib().rrc_synth_f32(device, seed, first_index, _ptr(dst), n_floats, stream))


# ------------------------------------------------------------------ FIR ---
def fir_plan(ntaps: int, deci: int, in_len: int, out_free: int):
    """(consume, need, out_n, wait_need, wait_on_output) — src/fir.rs:496-525."""
    v = [_sz(0) for _ in range(4)]
    w = _i(0)
    _ck(lib().rrc_fir_plan(ntaps, deci, in_len, out_free, *[C.byref(x) for x in v], C.byref(w)))
    return v[0].value, v[1].value, v[2].value, v[3].value, w.value


class Fir:
    def __init__(self, taps, deci: int = 1, device: int = 0, flags: int = 0, cplx: bool | None = None):
        taps = np.asarray(taps)
        self.cplx = bool(np.iscomplexobj(taps)) if cplx is None else cplx
        t = np.ascontiguousarray(taps, np.complex64 if self.cplx else np.float32)
        self.device, self.ntaps, self.deci = device, len(t), deci
        self.elem = 8 if self.cplx else 4
        h = C.c_void_p()
        fn = lib().rrc_fir_c32_create if self.cplx else lib().rrc_fir_f32_create
        _ck(fn(device, t.ctypes.data if len(t) else None, len(t), deci, flags, C.byref(h)))
        self.h = h.value

    def set_epilogue(self, kind: int, val: complex = 0):
        """Fuse the following MultiplyConst / AddConst / ComplexToMag2 into the output store (EPI_*)."""
        v = complex(val)
        _ck(lib().rrc_fir_set_epilogue(self.h, kind, v.real, v.imag))

    def set_translate(self, samp_rate: float, freq: float):
        _ck(lib().rrc_fir_set_translate(self.h, samp_rate, freq))

    def reset(self):
        _ck(lib().rrc_fir_reset(self.h))

    def set_input_u8iq(self, on: bool = True):
        """Inputs become u8 I/Q pairs, decoded like RtlSdrDecode inside the kernel's first load."""
        _ck(lib().rrc_fir_set_input_u8iq(self.h, int(on)))
        self.in_u8 = bool(on)

    @property
    def uses_real_taps(self) -> bool:
        y = _i(0)
        _ck(lib().rrc_fir_uses_real_taps(self.h, C.byref(y)))
        return bool(y.value)

    @property
    def uses_tensor_cores(self) -> bool:
        """True when runs go through the block-scaled fp16x3 tensor-core Toeplitz kernel (declared, like the real-tap path)."""
        y = _i(0)
        _ck(lib().rrc_fir_uses_tensor_cores(self.h, C.byref(y)))
        return bool(y.value)

    @property
    def kernel_name(self) -> str:
        buf = C.create_string_buffer(320)
        _ck(lib().rrc_fir_kernel_name(self.h, buf, 320))
        return buf.value.decode()

    def out_count(self, n_in: int) -> int:
        return 0 if n_in < self.ntaps + self.deci - 1 else (n_in - self.ntaps + 1) // self.deci

    def run(self, d_in, need: int, d_out, out_n: int, stream: int = 0):
        _ck(lib().rrc_fir_run(self.h, _ptr(d_in), need, _ptr(d_out), out_n, stream))

    def run_batch(self, d_in, in_stride, need, d_out, out_stride, out_n, nchan, stream: int = 0):
        _ck(lib().rrc_fir_run_batch(self.h, _ptr(d_in), in_stride, need, _ptr(d_out), out_stride, out_n, nchan, stream))

    def demod_run_batch(self, d_in, in_stride, need, gain, d_out, out_stride, out_n, nchan, stream: int = 0):
        _ck(lib().rrc_fir_c32_demod_run_batch(self.h, _ptr(d_in), in_stride, need, gain, _ptr(d_out), out_stride,
                                              out_n, nchan, stream))

    def demod_run_host_batch(self, x, n_in: int, nchan: int, gain: float, out) -> int:
        """Host buffers through the fused FIR + demod channelizer; returns outputs per channel."""
        xa = x.array if isinstance(x, PinnedBuffer) else x
        oa = out.array if isinstance(out, PinnedBuffer) else out
        per = _sz(0)
        _ck(lib().rrc_fir_c32_demod_run_host_batch(self.h, xa.ctypes.data, n_in, nchan, gain, oa.ctypes.data, len(oa) // nchan, C.byref(per)))
        return per.value

    def run_host(self, x, out=None) -> np.ndarray:
        """Whole-stream host->host (pipelined H2D/kernel/D2H).  In u8 I/Q mode x is the byte array."""
        dt = np.complex64 if self.cplx else np.float32
        u8 = getattr(self, "in_u8", False)
        xa = x.array if isinstance(x, PinnedBuffer) else np.ascontiguousarray(x, np.uint8 if u8 else dt)
        n_in = len(xa) // 2 if u8 else len(xa)
        n_out = self.out_count(n_in)
        oa = out.array if isinstance(out, PinnedBuffer) else (out if out is not None else np.empty(n_out, dt))
        n = _sz(0)
        _ck(lib().rrc_fir_run_host(self.h, xa.ctypes.data, n_in, oa.ctypes.data, C.byref(n)))
        return oa[: n.value]

    # convenience for tests: upload, run, download
    def filter(self, x: np.ndarray) -> np.ndarray:
        dt = np.complex64 if self.cplx else np.float32
        x = np.ascontiguousarray(x, dt)
        n_out = self.out_count(len(x))
        if n_out == 0:
            return np.empty(0, dt)
        need = (n_out - 1) * self.deci + self.ntaps
        din = DeviceBuffer.from_numpy(x[:need], self.device)
        dout = DeviceBuffer(n_out * self.elem, self.device)
        self.run(din, need, dout, n_out)
        return dout.download(dt, n_out)

    def __del__(self):
        try:
            lib().rrc_fir_destroy(self.h)
        except Exception:
            pass


# ----------------------------------------------------------- FFT filter ---
def fftfilt_ref_fft_size(ntaps: int):
    f, s = _sz(0), _sz(0)
    _ck(lib().rrc_fftfilt_ref_fft_size(ntaps, C.byref(f), C.byref(s)))
    return f.value, s.value


def fftfilt_plan(ntaps: int, buffered: int, in_len: int, out_free: int):
    """(blocks, consume, buffered_after, wait_need, wait_on_output) — src/fft_filter.rs:293-327."""
    v = [_sz(0) for _ in range(4)]
    w = _i(0)
    _ck(lib().rrc_fftfilt_plan(ntaps, buffered, in_len, out_free, *[C.byref(x) for x in v], C.byref(w)))
    return v[0].value, v[1].value, v[2].value, v[3].value, w.value


class FftFilt:
    def __init__(self, taps, device: int = 0, real: bool = False):
        """real=True: FftFilterFloat — f32 taps on an f32 stream (rrc_fftfilt_f32_create)."""
        self.real = real
        self.dtype = np.float32 if real else np.complex64
        t = np.ascontiguousarray(taps, self.dtype)
        self.device, self.ntaps = device, len(t)
        h = C.c_void_p()
        fn = lib().rrc_fftfilt_f32_create if real else lib().rrc_fftfilt_c32_create
        _ck(fn(device, t.ctypes.data if len(t) else None, len(t), C.byref(h)))
        self.h = h.value
        self.ref_fft_size, self.nsamples = fftfilt_ref_fft_size(len(t))

    def geometry(self):
        f, v = _sz(0), _sz(0)
        _ck(lib().rrc_fftfilt_geometry(self.h, C.byref(f), C.byref(v)))
        return f.value, v.value

    def reset(self, stream: int = 0):
        _ck(lib().rrc_fftfilt_reset(self.h, stream))

    def set_input_u8iq(self, on: bool = True):
        """Inputs become u8 I/Q pairs, decoded like RtlSdrDecode inside the kernel's first load."""
        _ck(lib().rrc_fftfilt_set_input_u8iq(self.h, int(on)))
        self.in_u8 = bool(on)

    def set_history(self, d_hist, n: int, stream: int = 0):
        _ck(lib().rrc_fftfilt_set_history(self.h, _ptr(d_hist), n, stream))

    def set_epilogue(self, kind: int, val: complex = 0):
        """Fuse the following MultiplyConst / AddConst / ComplexToMag2 into the output store (EPI_*)."""
        v = complex(val)
        _ck(lib().rrc_fftfilt_set_epilogue(self.h, kind, v.real, v.imag))

    def set_history_ptr(self, d_hist, n: int):
        """One-shot: the next run reads its left halo through this (possibly peer-mapped) device pointer."""
        _ck(lib().rrc_fftfilt_set_history_ptr(self.h, _ptr(d_hist), n))

    def run(self, d_in, n: int, d_out, stream: int = 0):
        _ck(lib().rrc_fftfilt_run(self.h, _ptr(d_in), n, _ptr(d_out), stream))

    def decim_run(self, d_in, n: int, deci: int, skip: int, d_out, stream: int = 0) -> int:
        no = _sz(0)
        _ck(lib().rrc_fftfilt_decim_run(self.h, _ptr(d_in), n, deci, skip, _ptr(d_out), C.byref(no), stream))
        return no.value

    def run_host(self, x, out=None) -> np.ndarray:
        u8 = getattr(self, "in_u8", False)
        xa = x.array if isinstance(x, PinnedBuffer) else np.ascontiguousarray(x, np.uint8 if u8 else self.dtype)
        n_in = len(xa) // 2 if u8 else len(xa)
        n_out = (n_in // self.nsamples) * self.nsamples
        oa = out.array if isinstance(out, PinnedBuffer) else (out if out is not None else np.empty(n_out, self.dtype))
        n = _sz(0)
        _ck(lib().rrc_fftfilt_run_host(self.h, xa.ctypes.data, n_in, oa.ctypes.data, C.byref(n)))
        return oa[: n.value]

    def decim_run_host(self, x, deci: int, out=None) -> np.ndarray:
        u8 = getattr(self, "in_u8", False)
        xa = x.array if isinstance(x, PinnedBuffer) else np.ascontiguousarray(x, np.uint8 if u8 else np.complex64)
        n_in = len(xa) // 2 if u8 else len(xa)
        n_out = ((n_in // self.nsamples) * self.nsamples + deci - 1) // deci
        oa = out.array if isinstance(out, PinnedBuffer) else (out if out is not None else np.empty(n_out, np.complex64))
        n = _sz(0)
        _ck(lib().rrc_fftfilt_decim_run_host(self.h, xa.ctypes.data, n_in, deci, oa.ctypes.data, C.byref(n)))
        return oa[: n.value]

    def filter(self, x: np.ndarray) -> np.ndarray:
        """Upload n samples, produce n outputs of the running convolution (stateful)."""
        x = np.ascontiguousarray(x, self.dtype)
        if len(x) == 0:
            return np.empty(0, self.dtype)
        din = DeviceBuffer.from_numpy(x, self.device)
        dout = DeviceBuffer(x.nbytes, self.device)
        self.run(din, len(x), dout)
        return dout.download(self.dtype, len(x))

    def __del__(self):
        try:
            lib().rrc_fftfilt_destroy(self.h)
        except Exception:
            pass


# ------------------------------------------------------------ resampler ---
class Resampler:
    def __init__(self, elem_size: int, interp: int, deci: int, device: int = 0):
        self.device, self.elem = device, elem_size
        h = C.c_void_p()
        _ck(lib().rrc_resampler_create(device, elem_size, interp, deci, C.byref(h)))
        self.h = h.value

    def state(self):
        i, d, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        p = _i(0)
        _ck(lib().rrc_resampler_state(self.h, C.byref(i), C.byref(d), C.byref(c), C.byref(p)))
        return i.value, d.value, c.value, bool(p.value)

    def reset(self):
        _ck(lib().rrc_resampler_reset(self.h))

    def set_state(self, counter: int, pending: np.ndarray | None = None):
        """Carried state of work() (src/rational_resampler.rs:101-105); see rrc_resampler_set_state."""
        pa = None if pending is None else np.ascontiguousarray(pending)
        assert pa is None or pa.nbytes == self.elem
        _ck(lib().rrc_resampler_set_state(self.h, counter, None if pa is None else pa.ctypes.data))

    def run(self, d_in, n_in: int, d_out, out_cap: int, stream: int = 0):
        """(consumed, produced, wait_on_output)"""
        c, p = _sz(0), _sz(0)
        w = _i(0)
        _ck(lib().rrc_resampler_run(self.h, _ptr(d_in), n_in, _ptr(d_out), out_cap, C.byref(c), C.byref(p), C.byref(w), stream))
        return c.value, p.value, w.value

    def run_host(self, x: np.ndarray, out_cap: int):
        xa = x.array if isinstance(x, PinnedBuffer) else np.ascontiguousarray(x)
        assert xa.dtype.itemsize == self.elem
        out = np.empty(out_cap, xa.dtype)
        c, p = _sz(0), _sz(0)
        _ck(lib().rrc_resampler_run_host(self.h, xa.ctypes.data, len(xa), out.ctypes.data, out_cap, C.byref(c), C.byref(p)))
        return c.value, out[: p.value]

    def run_host_into(self, xin: np.ndarray, xout: np.ndarray):
        """run_host on caller-owned (pinned) arrays: (consumed, produced)."""
        c, p = _sz(0), _sz(0)
        _ck(lib().rrc_resampler_run_host(self.h, xin.ctypes.data, len(xin), xout.ctypes.data, len(xout), C.byref(c), C.byref(p)))
        return c.value, p.value

    def work(self, x: np.ndarray, out_cap: int):
        """One work() call on host arrays via device buffers: (wait_on_output, consumed, out)."""
        x = np.ascontiguousarray(x)
        assert x.dtype.itemsize == self.elem
        din = DeviceBuffer.from_numpy(x, self.device)
        dout = DeviceBuffer(max(out_cap, 1) * self.elem, self.device)
        c, p, w = self.run(din, len(x), dout, out_cap)
        return w, c, dout.download(x.dtype, p)

    def __del__(self):
        try:
            lib().rrc_resampler_destroy(self.h)
        except Exception:
            pass


# ------------------------------------------------------------ FFT frames ---
def fftstream_plan(size: int, in_len: int, out_free: int):
    """(len, wait_need, wait_on_output) — src/fft_stream.rs:73-84."""
    ln, need, w = _sz(0), _sz(0), _i(0)
    _ck(lib().rrc_fftstream_plan(size, in_len, out_free, C.byref(ln), C.byref(need), C.byref(w)))
    return ln.value, need.value, w.value


class Fft:
    """Forward FFT of consecutive frames (FftStream / Fft compute)."""

    def __init__(self, size: int, device: int = 0):
        self.size, self.device = size, device
        h = C.c_void_p()
        _ck(lib().rrc_fft_c32_create(device, size, C.byref(h)))
        self.h = h.value

    def run(self, d_in, nframes: int, d_out, stream: int = 0):
        _ck(lib().rrc_fft_run(self.h, _ptr(d_in), nframes, _ptr(d_out), stream))

    def run_host(self, x, out=None) -> np.ndarray:
        xa = x.array if isinstance(x, PinnedBuffer) else np.ascontiguousarray(x, np.complex64)
        n_out = (len(xa) // self.size) * self.size
        oa = out.array if isinstance(out, PinnedBuffer) else (out if out is not None else np.empty(n_out, np.complex64))
        n = _sz(0)
        _ck(lib().rrc_fft_run_host(self.h, xa.ctypes.data, len(xa), oa.ctypes.data, C.byref(n)))
        return oa[: n.value]

    def transform(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.complex64)
        nf = len(x) // self.size
        if nf == 0:
            return np.empty(0, np.complex64)
        din = DeviceBuffer.from_numpy(x[: nf * self.size], self.device)
        dout = DeviceBuffer(nf * self.size * 8, self.device)
        self.run(din, nf, dout)
        return dout.download(np.complex64, nf * self.size)

    def __del__(self):
        try:
            lib().rrc_fft_destroy(self.h)
        except Exception:
            pass


# -------------------------------------------------------- rtlsdr decode ---
def rtlsdr_decode_plan(in_len_bytes: int, out_free: int):
    """(consume_bytes, produce, wait_need, wait_on_output) — src/rtlsdr_decode.rs:20-46."""
    v = [_sz(0) for _ in range(3)]
    w = _i(0)
    _ck(lib().rrc_rtlsdr_decode_plan(in_len_bytes, out_free, *[C.byref(x) for x in v], C.byref(w)))
    return v[0].value, v[1].value, v[2].value, w.value


def rtlsdr_decode(d_in, n_bytes: int, d_out, device: int = 0, stream: int = 0):
    _ck(lib().rrc_rtlsdr_decode_run(device, _ptr(d_in), n_bytes, _ptr(d_out), stream))


def rtlsdr_encode_plan(in_len: int, out_free_bytes: int):
    """(consume, produce_bytes, wait_need, wait_on_output) — src/rtlsdr_encode.rs:30-51."""
    v = [_sz(0) for _ in range(3)]
    w = _i(0)
    _ck(lib().rrc_rtlsdr_encode_plan(in_len, out_free_bytes, *[C.byref(x) for x in v], C.byref(w)))
    return v[0].value, v[1].value, v[2].value, w.value


def rtlsdr_encode(d_in, n: int, d_out, device: int = 0, stream: int = 0):
    _ck(lib().rrc_rtlsdr_encode_run(device, _ptr(d_in), n, _ptr(d_out), stream))


def rtlsdr_encode_host(x: np.ndarray, device: int = 0) -> np.ndarray:
    """RtlSdrEncode over a host buffer (pipelined H2D / kernel / D2H): c32 -> u8 I/Q bytes."""
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty(2 * len(x), np.uint8)
    n = _sz(0)
    _ck(lib().rrc_rtlsdr_encode_run_host(device, x.ctypes.data if len(x) else None, len(x), out.ctypes.data if len(out) else None, C.byref(n)))
    assert n.value == len(out)
    return out


def rtlsdr_decode_host(raw: np.ndarray, device: int = 0) -> np.ndarray:
    raw = np.ascontiguousarray(raw, np.uint8)
    out = np.empty(len(raw) // 2, np.complex64)
    n = _sz(0)
    _ck(lib().rrc_rtlsdr_decode_run_host(device, raw.ctypes.data if len(raw) else None, len(raw), out.ctypes.data if len(out) else None, C.byref(n)))
    return out[: n.value]


# ---------------------------------------------------------------- demod ---
def quad_demod(d_in, n_in: int, gain: float, d_out, device: int = 0, stream: int = 0):
    _ck(lib().rrc_quad_demod_run(device, _ptr(d_in), n_in, gain, _ptr(d_out), stream))


def quad_demod_host(x: np.ndarray, gain: float = 1.0, device: int = 0) -> np.ndarray:
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty(max(len(x) - 1, 0), np.float32)
    _ck(lib().rrc_quad_demod_run_host(device, x.ctypes.data, len(x), gain, out.ctypes.data))
    return out


# -------------------------------------------------------------- Hilbert ---
WINDOW_HAMMING, WINDOW_BLACKMAN, WINDOW_BLACKMAN_HARRIS, WINDOW_HAMMING_PARM = 0, 1, 2, 3


def make_window(window_type: int, ntaps: int, parm: float = 0.0) -> np.ndarray:
    """WindowType::make_window (src/window.rs:63-86), computed by the library on the host."""
    w = np.empty(ntaps, np.float32)
    _ck(lib().rrc_make_window(window_type, parm, ntaps, w.ctypes.data if ntaps else None))
    return w


def hilbert_taps(window: np.ndarray) -> np.ndarray:
    """fir::hilbert (src/fir.rs:660-680)."""
    w = np.ascontiguousarray(window, np.float32)
    t = np.empty(len(w), np.float32)
    _ck(lib().rrc_hilbert_taps(w.ctypes.data if len(w) else None, len(w), t.ctypes.data if len(w) else None))
    return t


class Hilbert:
    """Hilbert compute (src/hilbert.rs:86-125): f32 in, Complex out, ntaps of carried history."""

    def __init__(self, ntaps: int, window_type: int = WINDOW_HAMMING, parm: float = 0.0, device: int = 0, taps=None):
        self.device = device
        t = np.ascontiguousarray(taps, np.float32) if taps is not None else hilbert_taps(make_window(window_type, ntaps, parm))
        self.taps, self.ntaps = t, len(t)
        h = C.c_void_p()
        _ck(lib().rrc_hilbert_create(device, t.ctypes.data if len(t) else None, len(t), C.byref(h)))
        self.h = h.value

    def reset(self, stream: int = 0):
        _ck(lib().rrc_hilbert_reset(self.h, stream))

    def run(self, d_in, n: int, d_out, stream: int = 0):
        _ck(lib().rrc_hilbert_run(self.h, _ptr(d_in), n, _ptr(d_out), stream))

    def process(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        if len(x) == 0:
            return np.empty(0, np.complex64)
        din = DeviceBuffer.from_numpy(x, self.device)
        dout = DeviceBuffer(len(x) * 8, self.device)
        self.run(din, len(x), dout)
        return dout.download(np.complex64, len(x))

    def __del__(self):
        try:
            lib().rrc_hilbert_destroy(self.h)
        except Exception:
            pass


# ------------------------------------------------- sample-wise neighbours ---
def _map(fn, x, in_dt, out_dt, *vals, device: int = 0, offset: int = 0):
    """Upload, run one map, download.  offset (samples) misaligns both device pointers (ring windows do)."""
    x = np.ascontiguousarray(x, in_dt)
    n = len(x)
    if n == 0:
        return np.empty(0, out_dt)
    isz, osz = np.dtype(in_dt).itemsize, np.dtype(out_dt).itemsize
    din = DeviceBuffer((n + offset) * isz, device)
    din.upload(x, offset * isz)
    dout = DeviceBuffer((n + offset) * osz, device)
    _ck(fn(device, din.ptr + offset * isz, n, *vals, dout.ptr + offset * osz, 0))
    return dout.download(out_dt, n, offset * osz)


def multiply_const(x: np.ndarray, val, device: int = 0, offset: int = 0) -> np.ndarray:
    if np.iscomplexobj(x):
        v = complex(val)
        return _map(lib().rrc_multiply_const_c32_run, x, np.complex64, np.complex64, v.real, v.imag, device=device, offset=offset)
    return _map(lib().rrc_multiply_const_f32_run, x, np.float32, np.float32, float(val), device=device, offset=offset)


def add_const(x: np.ndarray, val, device: int = 0, offset: int = 0) -> np.ndarray:
    if np.iscomplexobj(x):
        v = complex(val)
        return _map(lib().rrc_add_const_c32_run, x, np.complex64, np.complex64, v.real, v.imag, device=device, offset=offset)
    return _map(lib().rrc_add_const_f32_run, x, np.float32, np.float32, float(val), device=device, offset=offset)


def complex_to_mag2(x: np.ndarray, device: int = 0, offset: int = 0) -> np.ndarray:
    return _map(lib().rrc_complex_to_mag2_run, x, np.complex64, np.float32, device=device, offset=offset)


def tee(x: np.ndarray, device: int = 0, offset_bytes: int = 0):
    x = np.ascontiguousarray(x)
    nb = x.nbytes
    din = DeviceBuffer(nb + offset_bytes + 16, device)
    din.upload(x.view(np.uint8), offset_bytes)
    o1, o2 = DeviceBuffer(nb + offset_bytes + 16, device), DeviceBuffer(nb + offset_bytes + 16, device)
    _ck(lib().rrc_tee_run(device, din.ptr + offset_bytes, nb, o1.ptr + offset_bytes, o2.ptr + offset_bytes, 0))
    return (o1.download(np.uint8, nb, offset_bytes).view(x.dtype), o2.download(np.uint8, nb, offset_bytes).view(x.dtype))


def iq_balance_alpha_from_tau(sample_rate: int, tau_seconds: float = 0.2) -> float:
    a = _f(0)
    _ck(lib().rrc_iq_balance_alpha_from_tau(sample_rate, tau_seconds, C.byref(a)))
    return a.value


class IqBalance:
    """IqBalance compute (src/iq_balance.rs:75-80) with the carried mean."""

    def __init__(self, alpha: float, device: int = 0):
        self.device = device
        h = C.c_void_p()
        _ck(lib().rrc_iq_balance_create(device, alpha, C.byref(h)))
        self.h = h.value

    def reset(self, stream: int = 0):
        _ck(lib().rrc_iq_balance_reset(self.h, stream))

    @property
    def mean(self) -> complex:
        m = np.zeros(2, np.float32)
        _ck(lib().rrc_iq_balance_mean(self.h, m.ctypes.data, 0))
        return complex(m[0], m[1])

    def run(self, d_in, n: int, d_out, stream: int = 0):
        _ck(lib().rrc_iq_balance_run(self.h, _ptr(d_in), n, _ptr(d_out), stream))

    def process(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.complex64)
        if len(x) == 0:
            return np.empty(0, np.complex64)
        din = DeviceBuffer.from_numpy(x, self.device)
        dout = DeviceBuffer(len(x) * 8, self.device)
        self.run(din, len(x), dout)
        return dout.download(np.complex64, len(x))

    def __del__(self):
        try:
            lib().rrc_iq_balance_destroy(self.h)
        except Exception:
            pass
