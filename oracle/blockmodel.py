"""Pure-Python restatement of rustradio's stream + Block::work() contract.

TEST INFRASTRUCTURE ONLY (see oracle/rr_oracle.c header).  Small cases only.

Restates, with the reference line ranges it follows (paths relative to
/root/reference):
  Tag / TagValue                   src/stream.rs:17-93
  Stream ring (produce/consume,    src/nowasm/circular_buffer.rs:174-216,
   read_buf tag re-basing)          :472-616
  ReadStream::eof                  src/stream.rs:237-246
  VectorSource::work               src/vector_source.rs:97-144
  FirFilter::work                  src/fir.rs:488-551
  FftFilter::work                  src/fft_filter.rs:289-355
  FftFilterFloat::work             src/fft_filter.rs:428-490
  RationalResampler::work / eof    src/rational_resampler.rs:154-213
  QuadratureDemod::work            src/quadrature_demod.rs:45-114
The sample arithmetic is delegated to oracle.oracle (the C restatement).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import Any

import numpy as np

from . import oracle as O

DEFAULT_STREAM_SIZE = 4_096_000  # bytes, src/stream.rs:105

# BlockRet kinds (src/block.rs:12-70)
AGAIN, PENDING, WAIT, EOF = "Again", "Pending", "WaitForStream", "EOF"


@dataclass(frozen=True)
class Tag:
    pos: int
    key: str
    val: Any  # ("Bool", True) / ("U64", 3) / ("String", "x") / ("Float", 1.0) / ("I64", -1)

    def at(self, pos: int) -> "Tag":
        return Tag(pos, self.key, self.val)


def tag_bool(pos, key, v):
    return Tag(pos, key, ("Bool", bool(v)))


def tag_u64(pos, key, v):
    return Tag(pos, key, ("U64", int(v)))


_stream_ids = itertools.count(1)


class Stream:
    """One circular buffer shared by a writer and a reader side."""

    def __init__(self, dtype, size_bytes: int = DEFAULT_STREAM_SIZE):
        self.dtype = np.dtype(dtype)
        self.cap = size_bytes // self.dtype.itemsize
        self.buf = np.zeros(2 * self.cap, self.dtype)  # emulates the double mapping
        self.rpos = self.wpos = self.used = 0
        self.tags: dict[int, list[Tag]] = {}
        self.id = next(_stream_ids)
        self.writer_alive = True
        self.reader_alive = True

    # -- writer side ---------------------------------------------------
    def free(self) -> int:
        return self.cap - self.used

    def write_buf(self) -> np.ndarray:
        """All free space, contiguous (circular_buffer.rs:607-615)."""
        return self.buf[self.wpos:self.wpos + self.free()]

    def produce(self, n: int, tags=()):
        """circular_buffer.rs:518-557."""
        for t in tags:
            assert t.pos < n, f"tag {t} out of range {n}"
        if n == 0:
            return
        assert self.free() >= n
        # mirror what the double mapping does for free
        lo = self.wpos
        seg = self.buf[lo:lo + n].copy()
        idx = (np.arange(lo, lo + n) % self.cap)
        self.buf[idx] = seg
        self.buf[idx + self.cap] = seg
        for t in tags:
            pos = (t.pos + self.wpos) % self.cap
            self.tags.setdefault(pos, []).append(t.at(pos))
        self.wpos = (self.wpos + n) % self.cap
        self.used += n

    # -- reader side ---------------------------------------------------
    def read_buf(self):
        """circular_buffer.rs:572-604: window + tags re-based, sorted by pos (stable)."""
        start, end = self.rpos, self.rpos + self.used
        out = []
        for n in sorted(self.tags):
            m = n % self.cap
            if end < self.cap and start < self.cap:
                if m < start or m >= end:
                    continue
            else:
                if m >= (end % self.cap) and m < start:
                    continue
            for t in self.tags[n]:
                out.append(t.at((t.pos + self.cap - start) % self.cap))
        out.sort(key=lambda t: t.pos)
        return self.buf[start:end], out

    def consume(self, n: int):
        """circular_buffer.rs:472-513."""
        if n == 0:
            return
        assert n <= self.used
        newpos = (self.rpos + n) % self.cap
        if newpos > self.rpos:
            keys = [k for k in self.tags if self.rpos <= k < newpos]
        else:
            keys = [k for k in self.tags if k >= self.rpos or k < newpos]
        for k in keys:
            del self.tags[k]
        self.rpos = newpos
        self.used -= n

    def eof(self) -> bool:
        """ReadStream::eof, src/stream.rs:237-246."""
        return (not self.writer_alive) and self.used == 0

    def closed_for_reader(self) -> bool:
        return not self.writer_alive


class BlockRet:
    def __init__(self, kind, stream=None, need=0):
        self.kind, self.stream, self.need = kind, stream, need

    def __repr__(self):
        return f"{self.kind}({self.stream.id if self.stream else ''},{self.need})"


class VectorSource:
    def __init__(self, data, repeat: int = 1, dtype=None, stream_bytes=DEFAULT_STREAM_SIZE):
        self.data = np.asarray(data, dtype=dtype)
        self.out = Stream(self.data.dtype, stream_bytes)
        self.repeat_n, self.count, self.pos = repeat, 0, 0

    def work(self) -> BlockRet:
        if len(self.data) == 0 or self.count >= self.repeat_n:
            return BlockRet(EOF)
        tags = []
        if self.pos == 0:
            tags = [tag_bool(0, "VectorSource::start", True), tag_u64(0, "VectorSource::repeat", self.count)]
            if self.count == 0:
                tags.append(tag_bool(0, "VectorSource::first", True))
        w = self.out.write_buf()
        if len(w) == 0:
            return BlockRet(WAIT, self.out, 1)
        n = min(len(w), len(self.data) - self.pos)
        w[:n] = self.data[self.pos:self.pos + n]
        self.out.produce(n, tags)
        self.pos += n
        if self.pos == len(self.data):
            self.count += 1
            if not (self.count < self.repeat_n):
                self.out.writer_alive = False  # block dropped by the caller in the reference tests
                return BlockRet(EOF)
            self.pos = 0
        return BlockRet(AGAIN)


class FirFilter:
    def __init__(self, src: Stream, taps, deci: int = 1, translate=None, stream_bytes=DEFAULT_STREAM_SIZE):
        taps = np.asarray(taps)
        assert len(taps) > 0 and deci != 0
        self.cplx = np.iscomplexobj(taps) or src.dtype.kind == "c"
        self.taps = taps.astype(np.complex64 if self.cplx else np.float32)
        self.phase = self.step = None
        if translate is not None:
            assert self.cplx
            self.taps, self.phase, self.step = O.fir_new_translator(self.taps, translate[0], translate[1], deci)
        self.ntaps, self.deci, self.src = len(self.taps), deci, src
        self.out = Stream(src.dtype, stream_bytes)

    def work(self) -> BlockRet:
        inp, tags = self.src.read_buf()
        absolute_minimum = self.ntaps + self.deci - 1
        if len(inp) < absolute_minimum:
            return BlockRet(WAIT, self.src, absolute_minimum)
        n = self.deci * ((len(inp) - self.ntaps + 1) // self.deci)
        need = n + self.ntaps - 1
        out = self.out.write_buf()
        if len(out) < 1:
            return BlockRet(WAIT, self.out, 1)
        n = min(n, len(out) * self.deci)
        out_n = n // self.deci
        need = n + self.ntaps - 1
        y = O.fir(inp[:need], self.taps, self.deci)
        assert len(y) == out_n
        if self.phase is not None:
            self.phase = O.fir_translate_output(y, self.phase, self.step)
        out[:out_n] = y
        tags = [t for t in tags if t.pos < n]
        self.src.consume(n)
        if self.deci != 1:
            tags = [t.at(t.pos // self.deci) for t in tags]
        self.out.produce(out_n, tags)
        return BlockRet(AGAIN)


class FftFilter:
    def __init__(self, src: Stream, taps, stream_bytes=DEFAULT_STREAM_SIZE):
        self.eng = O.FftFilt(taps)
        self.nsamples, self.src = self.eng.nsamples, src
        self.out = Stream(np.complex64, stream_bytes)
        self.buf = np.zeros(0, np.complex64)
        self.tags: list[Tag] = []

    def work(self) -> BlockRet:
        while True:
            o = self.out.write_buf()
            if self.nsamples > len(o):
                return BlockRet(WAIT, self.out, self.nsamples)
            inp, tags = self.src.read_buf()
            add = min(len(inp), self.nsamples - len(self.buf))
            tag_offset = len(self.buf)
            self.buf = np.concatenate([self.buf, inp[:add]])
            self.tags += [t.at(t.pos + tag_offset) for t in tags if t.pos < add]
            self.src.consume(add)
            if len(self.buf) < self.nsamples:
                return BlockRet(WAIT, self.src, self.nsamples - len(self.buf))
            o[:self.nsamples] = self.eng.run(self.buf)
            self.out.produce(self.nsamples, self.tags)
            self.buf = np.zeros(0, np.complex64)
            self.tags = []


class FftFilterFloat:
    def __init__(self, src: Stream, taps, stream_bytes=DEFAULT_STREAM_SIZE):
        self.src = src
        self.inner_in = Stream(np.complex64, stream_bytes)
        self.complex = FftFilter(self.inner_in, np.asarray(taps, np.float32).astype(np.complex64), stream_bytes)
        self.inner_out = self.complex.out
        self.out = Stream(np.float32, stream_bytes)

    def work(self) -> BlockRet:
        outer_in, tags = self.src.read_buf()
        inner_to = self.inner_in.write_buf()
        n = min(len(outer_in), len(inner_to))
        inner_to[:n] = outer_in[:n].astype(np.complex64)
        self.inner_in.produce(n, [t for t in tags if t.pos < n])
        self.src.consume(n)
        ret = self.complex.work()
        inner_from, tags = self.inner_out.read_buf()
        outer_to = self.out.write_buf()
        n = min(len(inner_from), len(outer_to))
        if n == 0 and len(inner_from) != 0:
            return BlockRet(WAIT, self.out, 1)
        outer_to[:n] = inner_from[:n].real
        tags = [t for t in tags if t.pos < n]
        self.inner_out.consume(n)
        self.out.produce(n, tags)
        if ret.kind == WAIT:
            if ret.stream is self.inner_in:
                return BlockRet(WAIT, self.src, ret.need)
            return BlockRet(WAIT, self.out, ret.need)
        return ret


class RationalResampler:
    def __init__(self, src: Stream, interp: int, deci: int, stream_bytes=DEFAULT_STREAM_SIZE):
        self.r = O.Resampler(src.dtype.itemsize, interp, deci)  # raises ValueError on 0
        self.src = src
        self.out = Stream(src.dtype, stream_bytes)

    def work(self) -> BlockRet:
        o = self.out.write_buf()
        if len(o) == 0:
            return BlockRet(WAIT, self.out, 1)
        inp, _ = self.src.read_buf()
        # The C restatement performs the pending flush and the counter loop;
        # it reads the input window only after the pending flush, like the
        # reference (:161-179).
        raw_in = inp.view(_uint(inp.dtype))
        ret, consumed, y = self.r.work(raw_in, len(o))
        o[:len(y)] = y.view(inp.dtype)
        self.src.consume(consumed)
        self.out.produce(len(y), [])
        return BlockRet(WAIT, self.out if ret == O.Resampler.WAIT_DST else self.src, 1)

    def eof(self) -> bool:
        return (not self.r.has_pending) and self.src.eof()


class RtlSdrDecode:
    """src/rtlsdr_decode.rs:18-48: u8 pairs -> Complex; tags dropped; never returns Again."""

    def __init__(self, src: Stream, stream_bytes=DEFAULT_STREAM_SIZE):
        self.src = src
        self.out = Stream(np.complex64, stream_bytes)

    def work(self) -> BlockRet:
        while True:
            inp, _ = self.src.read_buf()
            isamples = len(inp) & ~1
            if isamples == 0:
                return BlockRet(WAIT, self.src, 2)
            out = self.out.write_buf()
            if len(out) == 0:
                return BlockRet(WAIT, self.out, 1)
            isamples = min(isamples, len(out) * 2)
            osamples = isamples // 2
            out[:osamples] = O.rtlsdr_decode(inp[:isamples])
            self.src.consume(isamples)
            self.out.produce(osamples, [])


def _uint(dt):
    return {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[np.dtype(dt).itemsize]


class QuadratureDemod:
    def __init__(self, src: Stream, gain: float, stream_bytes=DEFAULT_STREAM_SIZE):
        self.src, self.gain = src, gain
        self.out = Stream(np.float32, stream_bytes)

    def work(self) -> BlockRet:
        while True:
            inp, _ = self.src.read_buf()
            if len(inp) < 2:
                return BlockRet(WAIT, self.src, 2)
            out = self.out.write_buf()
            if len(out) == 0:
                return BlockRet(WAIT, self.out, 1)
            n1 = min(len(inp) - 1, len(out))
            out[:n1] = O.quad_demod(inp[:n1 + 1], self.gain)
            self.src.consume(n1)
            self.out.produce(n1, [])


class Hilbert:
    """Hilbert::new / work (src/hilbert.rs:35-60,72-128): f32 in, Complex out, identity tags, one
    pass per call then Again."""

    def __init__(self, src: Stream, ntaps: int, window_type: int = O.WINDOW_HAMMING, parm: float = 0.0,
                 stream_bytes=DEFAULT_STREAM_SIZE):
        self.eng = O.Hilbert(ntaps, window_type, parm)
        self.ntaps, self.src = ntaps, src
        self.out = Stream(np.complex64, stream_bytes)

    def work(self) -> BlockRet:
        i, tags = self.src.read_buf()
        if len(i) == 0:
            return BlockRet(WAIT, self.src, 1)
        o = self.out.write_buf()
        if len(o) == 0:
            return BlockRet(WAIT, self.out, 1)
        inout = min(len(i), len(o))
        n = (self.ntaps + inout) - self.ntaps          # len - self.ntaps, :86-87
        if n == 0:
            return BlockRet(WAIT, self.src if len(i) < len(o) else self.out, 1)
        o[:n] = self.eng.work(i[:inout])
        self.out.produce(n, [t for t in tags if t.pos < n])
        self.src.consume(n)
        return BlockRet(AGAIN)


class _Sync:
    """The macro-generated `sync` work loop (rustradio_macros_code/src/lib.rs:458-513): one input,
    any number of outputs, every output gets the input tags at unchanged positions."""

    out_dtypes: tuple = ()

    def __init__(self, src: Stream, out_dtypes, stream_bytes=DEFAULT_STREAM_SIZE):
        self.src = src
        self.outs = [Stream(dt, stream_bytes) for dt in out_dtypes]
        self.out = self.outs[0]

    def process(self, x):                                # -> tuple of arrays, one per output
        raise NotImplementedError

    def work(self) -> BlockRet:
        while True:
            i, tags = self.src.read_buf()
            if len(i) == 0:
                return BlockRet(WAIT, self.src, 1)
            n = len(i)
            ws = []
            for o in self.outs:
                w = o.write_buf()
                if len(w) == 0:
                    return BlockRet(WAIT, o, 1)
                ws.append(w)
            n = min([n] + [len(w) for w in ws])
            ys = self.process(i[:n])
            for w, y in zip(ws, ys):
                w[:n] = y
            keep = [t for t in tags if t.pos < n]
            self.src.consume(n)
            for o in self.outs:
                o.produce(n, keep)


class MultiplyConst(_Sync):
    def __init__(self, src: Stream, val, stream_bytes=DEFAULT_STREAM_SIZE):
        super().__init__(src, [src.dtype], stream_bytes)
        self.val = val

    def process(self, x):
        return (O.multiply_const(x, self.val),)


class AddConst(_Sync):
    def __init__(self, src: Stream, val, stream_bytes=DEFAULT_STREAM_SIZE):
        super().__init__(src, [src.dtype], stream_bytes)
        self.val = val

    def process(self, x):
        return (O.add_const(x, self.val),)


class ComplexToMag2(_Sync):
    def __init__(self, src: Stream, stream_bytes=DEFAULT_STREAM_SIZE):
        super().__init__(src, [np.float32], stream_bytes)

    def process(self, x):
        return (O.complex_to_mag2(x),)


class Tee(_Sync):
    """Tee::new(src) -> (Self, out1, out2) (src/tee.rs:9-24)."""

    def __init__(self, src: Stream, stream_bytes=DEFAULT_STREAM_SIZE):
        super().__init__(src, [src.dtype, src.dtype], stream_bytes)
        self.out1, self.out2 = self.outs

    def process(self, x):
        return (x, x)


class IqBalance(_Sync):
    """IqBalance::with_alpha (src/iq_balance.rs:62-80)."""

    def __init__(self, src: Stream, alpha: float, stream_bytes=DEFAULT_STREAM_SIZE):
        super().__init__(src, [np.complex64], stream_bytes)
        self.eng = O.IqBalance(alpha)

    def process(self, x):
        return (self.eng.work(x),)
