"""ctypes binding of the block-level contract (rrb_* in include/rustradio_cuda.h):
rustradio's Block / ReadStream / WriteStream / Tag surface over the CUDA kernels.
Test/bench harness only — the blocks themselves are C++ (csrc/blocks.cu)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any

import numpy as np

from .api import RrcError, lib, _ck

AGAIN, PENDING, WAIT, EOF = 0, 1, 2, 3
HOST, DEVICE, HOST_PINNED = 0, 1, 2
REPEAT_INFINITE = 2**64 - 1
SIGMF_TYPES = {np.dtype(np.complex64): "cf32", np.dtype(np.float32): "rf32", np.dtype(np.uint8): "ru8", np.dtype(np.int32): "ri32"}
DEFAULT_STREAM_SIZE = 4_096_000
_KINDS = ["String", "Float", "Bool", "U64", "I64"]


class _CTag(C.Structure):
    _fields_ = [("pos", C.c_uint64), ("key", C.c_char_p), ("kind", C.c_int), ("s", C.c_char_p), ("f", C.c_float),
                ("b", C.c_int), ("u", C.c_uint64), ("i", C.c_int64)]


@dataclass(frozen=True)
class Tag:
    pos: int
    key: str
    val: Any   # ("Bool", True) / ("U64", 3) / ("String", "x") / ("Float", 1.0) / ("I64", -1)


_sz, _vp, _i = C.c_size_t, C.c_void_p, C.c_int
_P = C.POINTER
_SIGS = {
    "rrb_stream_new": [_sz, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_wstream_write": [_vp, _vp, _sz, _P(_CTag), _sz, _P(_sz)],
    "rrb_wstream_free": [_vp, _P(_sz)],
    "rrb_wstream_id": [_vp, _P(_sz)],
    "rrb_wstream_drop": [_vp],
    "rrb_rstream_read": [_vp, _vp, _sz, _P(_sz), _P(_sz)],
    "rrb_rstream_tag": [_vp, _sz, _P(_CTag)],
    "rrb_rstream_consume": [_vp, _sz],
    "rrb_rstream_id": [_vp, _P(_sz)],
    "rrb_rstream_capacity": [_vp, _P(_sz)],
    "rrb_rstream_eof": [_vp, _P(_i)],
    "rrb_rstream_drop": [_vp],
    "rrb_vector_source_new": [_vp, _sz, _sz, C.c_uint64, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_file_source_new": [C.c_char_p, _sz, C.c_uint64, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_sigmf_source_new": [C.c_char_p, _sz, C.c_char_p, C.c_double, _i, C.c_uint64, _sz, _i, _i, _P(_vp), _P(_vp), _P(C.c_double), _P(_i)],
    "rrb_fir_filter_new": [_vp, _i, _vp, _sz, _sz, _i, C.c_float, C.c_float, C.c_uint, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_fft_filter_new": [_vp, _vp, _sz, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_fft_filter_float_new": [_vp, _vp, _sz, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_rational_resampler_new": [_vp, _sz, _sz, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_quadrature_demod_new": [_vp, C.c_float, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_rtlsdr_decode_new": [_vp, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_rtlsdr_encode_new": [_vp, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_file_sink_new": [_vp, C.c_char_p, _i, _i, _i, _P(_vp)],
    "rrb_fft_stream_new": [_vp, _sz, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_hilbert_new": [_vp, _sz, _i, C.c_float, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_multiply_const_new": [_vp, _i, C.c_float, C.c_float, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_add_const_new": [_vp, _i, C.c_float, C.c_float, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_complex_to_mag2_new": [_vp, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_iq_balance_new": [_vp, C.c_float, _sz, _i, _i, _P(_vp), _P(_vp)],
    "rrb_tee_new": [_vp, _sz, _i, _i, _P(_vp), _P(_vp), _P(_vp)],
    "rrb_block_work": [_vp, _P(_i), _P(_sz), _P(_sz)],
    "rrb_block_eof": [_vp, _P(_i)],
    "rrb_block_drop": [_vp],
    "rrb_graph_run": [_P(_vp), _sz],
}
_bound = False


def exported_symbols() -> list[str]:
    return sorted(_SIGS) + ["rrb_block_name"]


def _L():
    global _bound
    L = lib()
    if not _bound:
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        L.rrb_block_name.argtypes = [_vp]
        L.rrb_block_name.restype = C.c_char_p
        _bound = True
    return L


def _ctags(tags):
    arr = (_CTag * max(len(tags), 1))()
    keep = []
    for k, t in enumerate(tags):
        kind, v = t.val
        arr[k].pos = t.pos
        kb = t.key.encode(); keep.append(kb)
        arr[k].key = kb
        arr[k].kind = _KINDS.index(kind)
        if kind == "String":
            sb = str(v).encode(); keep.append(sb); arr[k].s = sb
        elif kind == "Float":
            arr[k].f = float(v)
        elif kind == "Bool":
            arr[k].b = int(bool(v))
        elif kind == "U64":
            arr[k].u = int(v)
        else:
            arr[k].i = int(v)
    return arr, keep


class ReadStream:
    def __init__(self, h, dtype):
        self.h, self.dtype = h, np.dtype(dtype)

    @property
    def id(self) -> int:
        v = _sz(0); _ck(_L().rrb_rstream_id(self.h, C.byref(v))); return v.value

    @property
    def capacity(self) -> int:
        v = _sz(0); _ck(_L().rrb_rstream_capacity(self.h, C.byref(v))); return v.value

    def read_buf(self, max_samples: int | None = None):
        """(samples, tags) of the current window, like ReadStream::read_buf (no consume)."""
        n, nt = _sz(0), _sz(0)
        _ck(_L().rrb_rstream_read(self.h, None, 0, C.byref(n), C.byref(nt)))
        m = n.value if max_samples is None else min(n.value, max_samples)
        out = np.empty(m, self.dtype)
        _ck(_L().rrb_rstream_read(self.h, out.ctypes.data, m, C.byref(n), C.byref(nt)))
        tags = []
        for k in range(nt.value):
            ct = _CTag()
            _ck(_L().rrb_rstream_tag(self.h, k, C.byref(ct)))
            kind = _KINDS[ct.kind]
            v = {"String": (ct.s or b"").decode(), "Float": ct.f, "Bool": bool(ct.b), "U64": ct.u, "I64": ct.i}[kind]
            tags.append(Tag(ct.pos, ct.key.decode(), (kind, v)))
        return out, tags

    def __len__(self):
        n = _sz(0); _ck(_L().rrb_rstream_read(self.h, None, 0, C.byref(n), None)); return n.value

    def consume(self, n: int):
        _ck(_L().rrb_rstream_consume(self.h, n))

    def eof(self) -> bool:
        v = _i(0); _ck(_L().rrb_rstream_eof(self.h, C.byref(v))); return bool(v.value)

    def _take(self):
        h, self.h = self.h, None
        return h

    def __del__(self):
        try:
            if self.h:
                _L().rrb_rstream_drop(self.h)
        except Exception:
            pass


class WriteStream:
    def __init__(self, h, dtype):
        self.h, self.dtype = h, np.dtype(dtype)

    @property
    def id(self) -> int:
        v = _sz(0); _ck(_L().rrb_wstream_id(self.h, C.byref(v))); return v.value

    def free(self) -> int:
        v = _sz(0); _ck(_L().rrb_wstream_free(self.h, C.byref(v))); return v.value

    def write(self, data, tags=()) -> int:
        a = np.ascontiguousarray(data, self.dtype)
        arr, keep = _ctags(list(tags))
        w = _sz(0)
        _ck(_L().rrb_wstream_write(self.h, a.ctypes.data, len(a), arr, len(tags), C.byref(w)))
        return w.value

    def drop(self):
        if self.h:
            _L().rrb_wstream_drop(self.h); self.h = None

    def __del__(self):
        try:
            self.drop()
        except Exception:
            pass


def new_stream(dtype, size_bytes: int = DEFAULT_STREAM_SIZE, residency: int = DEVICE, device: int = 0):
    w, r = _vp(), _vp()
    _ck(_L().rrb_stream_new(np.dtype(dtype).itemsize, size_bytes, residency, device, C.byref(w), C.byref(r)))
    return WriteStream(w.value, dtype), ReadStream(r.value, dtype)


@dataclass
class BlockRet:
    kind: int
    stream_id: int = 0
    need: int = 0


class Block:
    def __init__(self, h):
        self.h = h

    def work(self) -> BlockRet:
        k, sid, need = _i(0), _sz(0), _sz(0)
        _ck(_L().rrb_block_work(self.h, C.byref(k), C.byref(sid), C.byref(need)))
        return BlockRet(k.value, sid.value, need.value)

    def eof(self) -> bool:
        v = _i(0); _ck(_L().rrb_block_eof(self.h, C.byref(v))); return bool(v.value)

    @property
    def name(self) -> str:
        return _L().rrb_block_name(self.h).decode()

    def drop(self):
        if self.h:
            _L().rrb_block_drop(self.h); self.h = None

    def __del__(self):
        try:
            self.drop()
        except Exception:
            pass


def _mk(fn, out_dtype, *args):
    """Call a constructor.  A ReadStream among `args` is the block's input: the C side consumes it iff
    the call returns RRC_OK (include/rustradio_cuda.h), so the Python wrapper gives up its handle only then."""
    b, o = _vp(), _vp()
    _ck(fn(*[a.h if isinstance(a, ReadStream) else a for a in args], C.byref(b), C.byref(o)))
    for a in args:
        if isinstance(a, ReadStream):
            a._take()
    return Block(b.value), ReadStream(o.value, out_dtype)


def VectorSource(data, repeat: int = 1, size_bytes=DEFAULT_STREAM_SIZE, residency=DEVICE, device=0):
    a = np.ascontiguousarray(data)
    return _mk(_L().rrb_vector_source_new, a.dtype, a.ctypes.data if len(a) else None, len(a), a.dtype.itemsize, repeat,
               size_bytes, residency, device)


def FileSource(path, dtype, repeat: int = 1, size_bytes=DEFAULT_STREAM_SIZE, residency=DEVICE, device=0):
    """FileSource::<T>::builder(path).repeat(Repeat::finite(repeat)).build() (src/file_source.rs:11-40)."""
    dt = np.dtype(dtype)
    return _mk(_L().rrb_file_source_new, dt, str(path).encode(), dt.itemsize, repeat, size_bytes, residency, device)


def SigMFSource(path, dtype, sample_rate=None, ignore_type_error=False, repeat: int = 1, size_bytes=DEFAULT_STREAM_SIZE,
                residency=DEVICE, device=0):
    """SigMFSource::<T>::builder(path)...build() (src/sigmf.rs:229-268) -> (block, out, sample_rate or None)."""
    dt = np.dtype(dtype)
    b, o = _vp(), _vp()
    rate, has = C.c_double(0), _i(0)
    _ck(_L().rrb_sigmf_source_new(str(path).encode(), dt.itemsize, SIGMF_TYPES[dt].encode(), -1.0 if sample_rate is None else float(sample_rate),
                                  int(ignore_type_error), repeat, size_bytes, residency, device, C.byref(b), C.byref(o), C.byref(rate), C.byref(has)))
    return Block(b.value), ReadStream(o.value, dt), (rate.value if has.value else None)


def FirFilter(src: ReadStream, taps, deci: int = 1, translate=None, flags: int = 0, size_bytes=DEFAULT_STREAM_SIZE,
              residency=DEVICE, device=0):
    """FirFilter::builder(taps).deci(deci).translate(fs, f).build(src) -> (block, out)."""
    cplx = src.dtype == np.complex64
    t = np.ascontiguousarray(taps, np.complex64 if cplx else np.float32)
    tr = translate or (0.0, 0.0)
    return _mk(_L().rrb_fir_filter_new, src.dtype, src, int(cplx), t.ctypes.data if len(t) else None, len(t), deci,
               int(translate is not None), tr[0], tr[1], flags, size_bytes, residency, device)


def FftFilter(src: ReadStream, taps, size_bytes=DEFAULT_STREAM_SIZE, residency=DEVICE, device=0):
    t = np.ascontiguousarray(taps, np.complex64)
    return _mk(_L().rrb_fft_filter_new, np.complex64, src, t.ctypes.data if len(t) else None, len(t),
               size_bytes, residency, device)


def FftFilterFloat(src: ReadStream, taps, size_bytes=DEFAULT_STREAM_SIZE, residency=DEVICE, device=0):
    t = np.ascontiguousarray(taps, np.float32)
    return _mk(_L().rrb_fft_filter_float_new, np.float32, src, t.ctypes.data if len(t) else None, len(t),
               size_bytes, residency, device)


def RationalResampler(src: ReadStream, interp: int, deci: int, size_bytes=DEFAULT_STREAM_SIZE, residency=DEVICE, device=0):
    if interp == 0 or deci == 0:
        raise RrcError(-1, f"RationalResampler created using {'interp' if interp == 0 else 'deci'} 0")
    return _mk(_L().rrb_rational_resampler_new, src.dtype, src, interp, deci, size_bytes, residency, device)


def QuadratureDemod(src: ReadStream, gain: float, size_bytes=DEFAULT_STREAM_SIZE, residency=DEVICE, device=0):
    return _mk(_L().rrb_quadrature_demod_new, np.float32, src, gain, size_bytes, residency, device)


def FftStream(src: ReadStream, size: int, size_bytes=DEFAULT_STREAM_SIZE, residency=DEVICE, device=0):
    """FftStream::new(src, size) (src/fft_stream.rs:38-60)."""
    return _mk(_L().rrb_fft_stream_new, np.complex64, src, size, size_bytes, residency, device)


def RtlSdrDecode(src: ReadStream, size_bytes=DEFAULT_STREAM_SIZE, residency=DEVICE, device=0):
    """RtlSdrDecode::new(src) (src/rtlsdr_decode.rs:9-16): ReadStream<u8> -> ReadStream<Complex>."""
    return _mk(_L().rrb_rtlsdr_decode_new, np.complex64, src, size_bytes, residency, device)


def RtlSdrEncode(src: ReadStream, size_bytes=DEFAULT_STREAM_SIZE, residency=DEVICE, device=0):
    """RtlSdrEncode::new(src) (src/rtlsdr_encode.rs:12-20): ReadStream<Complex> -> ReadStream<u8>."""
    return _mk(_L().rrb_rtlsdr_encode_new, np.uint8, src, size_bytes, residency, device)


FILE_CREATE, FILE_OVERWRITE, FILE_APPEND = 0, 1, 2


def FileSink(src: ReadStream, path, mode: int = FILE_CREATE, flush: bool = False, device=0):
    """FileSink::<T>::builder(path).mode(mode).flush(flush).build(src) (src/file_sink.rs:24-115); a sink: returns the block only."""
    b = _vp()
    _ck(_L().rrb_file_sink_new(src.h, str(path).encode(), mode, int(flush), device, C.byref(b)))
    src._take()
    return Block(b.value)


def Hilbert(src: ReadStream, ntaps: int, window_type: int = 0, window_parm: float = 0.0, size_bytes=DEFAULT_STREAM_SIZE,
            residency=DEVICE, device=0):
    """Hilbert::new(src, ntaps, &window_type) (src/hilbert.rs:35-60): ReadStream<Float> -> ReadStream<Complex>."""
    return _mk(_L().rrb_hilbert_new, np.complex64, src, ntaps, window_type, window_parm, size_bytes, residency, device)


def MultiplyConst(src: ReadStream, val, size_bytes=DEFAULT_STREAM_SIZE, residency=DEVICE, device=0):
    """MultiplyConst::new(src, val) (src/multiply_const.rs:5-23)."""
    cplx = src.dtype == np.complex64
    v = complex(val)
    return _mk(_L().rrb_multiply_const_new, src.dtype, src, int(cplx), v.real, v.imag, size_bytes, residency, device)


def AddConst(src: ReadStream, val, size_bytes=DEFAULT_STREAM_SIZE, residency=DEVICE, device=0):
    """AddConst::new(src, val) (src/add_const.rs:24-44)."""
    cplx = src.dtype == np.complex64
    v = complex(val)
    return _mk(_L().rrb_add_const_new, src.dtype, src, int(cplx), v.real, v.imag, size_bytes, residency, device)


def ComplexToMag2(src: ReadStream, size_bytes=DEFAULT_STREAM_SIZE, residency=DEVICE, device=0):
    """ComplexToMag2::new(src) (src/complex_to_mag2.rs:7-20)."""
    return _mk(_L().rrb_complex_to_mag2_new, np.float32, src, size_bytes, residency, device)


def IqBalance(src: ReadStream, alpha: float, size_bytes=DEFAULT_STREAM_SIZE, residency=DEVICE, device=0):
    """IqBalance::with_alpha(src, alpha) (src/iq_balance.rs:62-73)."""
    return _mk(_L().rrb_iq_balance_new, np.complex64, src, alpha, size_bytes, residency, device)


def Tee(src: ReadStream, size_bytes=DEFAULT_STREAM_SIZE, residency=DEVICE, device=0):
    """Tee::new(src) -> (block, out1, out2) (src/tee.rs:9-18)."""
    b, o1, o2 = _vp(), _vp(), _vp()
    dt = src.dtype
    _ck(_L().rrb_tee_new(src.h, size_bytes, residency, device, C.byref(b), C.byref(o1), C.byref(o2)))
    src._take()
    return Block(b.value), ReadStream(o1.value, dt), ReadStream(o2.value, dt)


def graph_run(blocks):
    arr = (_vp * len(blocks))(*[b.h for b in blocks])
    _ck(_L().rrb_graph_run(arr, len(blocks)))
