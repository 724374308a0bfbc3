//! The four GPU blocks behind rustradio's `Block` trait (src/block.rs:115-126).
//!
//! This variant keeps rustradio's own host-resident streams (`ReadStream`/`WriteStream`,
//! src/stream.rs:180-327) and stages each `work()` window through device scratch, which needs no
//! change inside rustradio.  The zero-copy variant (device-resident double-mapped rings so chained
//! GPU blocks never round-trip through host memory) is what `csrc/blocks.cu` implements and needs
//! rustradio's `sys::Buffer` to become pluggable.
//!
//! UNCOMPILED in this repository's environment (no Rust toolchain).
use std::ffi::c_void;
use std::ptr;

use rustradio::block::{Block, BlockEOF, BlockName, BlockRet};
use rustradio::stream::{ReadStream, WriteStream};
use rustradio::{Complex, Float, Result};

use crate::ffi::{self, check};

struct Scratch { ptr: *mut c_void, cap: usize, dev: i32 }
impl Scratch {
    fn new(dev: i32) -> Self { Self { ptr: ptr::null_mut(), cap: 0, dev } }
    fn reserve(&mut self, bytes: usize) -> Result<*mut c_void> {
        if bytes > self.cap {
            // SAFETY: plain FFI; pointers come from the same allocator.
            unsafe {
                if !self.ptr.is_null() { check(ffi::rrc_free_device(self.dev, self.ptr))?; }
                check(ffi::rrc_malloc_device(self.dev, bytes, &mut self.ptr))?;
            }
            self.cap = bytes;
        }
        Ok(self.ptr)
    }
}
impl Drop for Scratch {
    fn drop(&mut self) { if !self.ptr.is_null() { unsafe { ffi::rrc_free_device(self.dev, self.ptr); } } }
}
// SAFETY: a block is driven by one thread at a time (src/mtgraph.rs:77-120); the handles are plain device resources.
unsafe impl Send for Scratch {}

/// GPU `FirFilter<Complex>`: same `new(src, taps)` / builder-style `with_deci` as src/fir.rs:357-386.
pub struct CudaFirFilter {
    h: *mut ffi::rrc_fir_t,
    ntaps: usize,
    deci: usize,
    dev: i32,
    sin: Scratch,
    sout: Scratch,
    src: ReadStream<Complex>,
    dst: WriteStream<Complex>,
}
unsafe impl Send for CudaFirFilter {}

impl CudaFirFilter {
    pub fn new(src: ReadStream<Complex>, taps: impl AsRef<[Complex]>) -> (Self, ReadStream<Complex>) {
        Self::with_deci(src, taps, 1)
    }
    pub fn with_deci(src: ReadStream<Complex>, taps: impl AsRef<[Complex]>, deci: usize) -> (Self, ReadStream<Complex>) {
        let taps = taps.as_ref();
        assert!(!taps.is_empty());      // src/fir.rs:372
        assert_ne!(deci, 0);            // src/fir.rs:319
        let mut h = ptr::null_mut();
        // SAFETY: Complex<f32> is repr(C) (re, im).
        unsafe { check(ffi::rrc_fir_c32_create(0, taps.as_ptr().cast(), taps.len(), deci, 0, &mut h)).expect("rrc_fir_c32_create"); }
        let (dst, dr) = rustradio::stream::new_stream();
        (Self { h, ntaps: taps.len(), deci, dev: 0, sin: Scratch::new(0), sout: Scratch::new(0), src, dst }, dr)
    }
}
impl Drop for CudaFirFilter { fn drop(&mut self) { unsafe { ffi::rrc_fir_destroy(self.h); } } }
impl BlockName for CudaFirFilter { fn block_name(&self) -> &str { "CudaFirFilter<Complex>" } }
impl BlockEOF for CudaFirFilter { fn eof(&mut self) -> bool { self.src.eof() } }

impl Block for CudaFirFilter {
    fn work(&mut self) -> Result<BlockRet<'_>> {
        let (input, mut tags) = self.src.read_buf()?;
        let mut out = self.dst.write_buf()?;
        let (mut n, mut need, mut out_n, mut wait_need, mut wait_out) = (0usize, 0usize, 0usize, 0usize, 0i32);
        // The integer part of FirFilter::work (src/fir.rs:496-525), shared with the C++ mirror.
        unsafe { check(ffi::rrc_fir_plan(self.ntaps, self.deci, input.len(), out.len(), &mut n, &mut need, &mut out_n, &mut wait_need, &mut wait_out))?; }
        if n == 0 {
            return Ok(if wait_out != 0 { BlockRet::WaitForStream(&self.dst, wait_need) } else { BlockRet::WaitForStream(&self.src, wait_need) });
        }
        let (din, dout) = (self.sin.reserve(need * 8)?, self.sout.reserve(out_n * 8)?);
        unsafe {
            check(ffi::rrc_memcpy_h2d(self.dev, din, input.slice().as_ptr().cast(), need * 8, ptr::null_mut()))?;
            check(ffi::rrc_fir_run(self.h, din, need, dout, out_n, ptr::null_mut()))?;
            check(ffi::rrc_memcpy_d2h(self.dev, out.slice().as_mut_ptr().cast(), dout, out_n * 8, ptr::null_mut()))?;
            check(ffi::rrc_stream_sync(self.dev, ptr::null_mut()))?;
        }
        tags.retain(|t| t.pos() < n);                       // src/fir.rs:536
        input.consume(n);
        if self.deci != 1 { for t in &mut tags { t.set_pos(t.pos() / self.deci); } }
        out.produce(out_n, &tags);
        Ok(BlockRet::Again)
    }
}

/// GPU `FftFilter`: `new(src, taps)` like src/fft_filter.rs:241-255; same count rule (whole blocks of
/// `nsamples`, partial block retained) via `rrc_fftfilt_plan`.
pub struct CudaFftFilter {
    h: *mut ffi::rrc_fftfilt_t,
    ntaps: usize,
    nsamples: usize,
    buf: Vec<Complex>,                  // self.buf of the reference (host side, < nsamples samples)
    tags: Vec<rustradio::stream::Tag>,
    dev: i32,
    sin: Scratch,
    sout: Scratch,
    src: ReadStream<Complex>,
    dst: WriteStream<Complex>,
}
unsafe impl Send for CudaFftFilter {}

impl CudaFftFilter {
    pub fn new<T: Into<Vec<Complex>>>(src: ReadStream<Complex>, taps: T) -> (Self, ReadStream<Complex>) {
        let taps: Vec<Complex> = taps.into();
        assert!(!taps.is_empty());
        let mut h = ptr::null_mut();
        let (mut fft_size, mut nsamples) = (0usize, 0usize);
        unsafe {
            check(ffi::rrc_fftfilt_c32_create(0, taps.as_ptr().cast(), taps.len(), &mut h)).expect("rrc_fftfilt_c32_create");
            check(ffi::rrc_fftfilt_ref_fft_size(taps.len(), &mut fft_size, &mut nsamples)).unwrap();
        }
        let (dst, dr) = rustradio::stream::new_stream();
        (Self { h, ntaps: taps.len(), nsamples, buf: Vec::with_capacity(nsamples), tags: Vec::new(), dev: 0,
                sin: Scratch::new(0), sout: Scratch::new(0), src, dst }, dr)
    }
}
impl Drop for CudaFftFilter { fn drop(&mut self) { unsafe { ffi::rrc_fftfilt_destroy(self.h); } } }
impl BlockName for CudaFftFilter { fn block_name(&self) -> &str { "CudaFftFilter" } }
impl BlockEOF for CudaFftFilter { fn eof(&mut self) -> bool { self.src.eof() } }

impl Block for CudaFftFilter {
    fn work(&mut self) -> Result<BlockRet<'_>> {
        use rustradio::stream::Tag;
        let mut o = self.dst.write_buf()?;
        let (input, tags) = self.src.read_buf()?;
        let (mut blocks, mut consume, mut after, mut wait_need, mut wait_out) = (0usize, 0usize, 0usize, 0usize, 0i32);
        unsafe { check(ffi::rrc_fftfilt_plan(self.ntaps, self.buf.len(), input.len(), o.len(), &mut blocks, &mut consume, &mut after, &mut wait_need, &mut wait_out))?; }
        let s = self.nsamples;
        let buffered = self.buf.len();
        // samples handed to the filter: (buf ++ input[..consume]); the first blocks*s of them are filtered now
        let mut staged: Vec<Complex> = Vec::with_capacity(buffered + consume);
        staged.extend_from_slice(&self.buf);
        staged.extend_from_slice(&input.slice()[..consume]);
        if blocks > 0 {
            let n = blocks * s;
            let (din, dout) = (self.sin.reserve(n * 8)?, self.sout.reserve(n * 8)?);
            unsafe {
                check(ffi::rrc_memcpy_h2d(self.dev, din, staged.as_ptr().cast(), n * 8, ptr::null_mut()))?;
                check(ffi::rrc_fftfilt_run(self.h, din.cast(), n, dout.cast(), ptr::null_mut()))?;
                check(ffi::rrc_memcpy_d2h(self.dev, o.slice().as_mut_ptr().cast(), dout, n * 8, ptr::null_mut()))?;
                check(ffi::rrc_stream_sync(self.dev, ptr::null_mut()))?;
            }
        }
        self.buf.clear();
        self.buf.extend_from_slice(&staged[blocks * s..]);
        debug_assert_eq!(self.buf.len(), after);
        // tags (src/fft_filter.rs:309-313): position in (buf ++ input), emitted with their block
        let mut all = std::mem::take(&mut self.tags);
        all.extend(tags.into_iter().filter(|t| t.pos() < consume).map(|t| Tag::new(t.pos() + buffered, t.key(), t.val().clone())));
        let (emit, keep): (Vec<_>, Vec<_>) = all.into_iter().partition(|t| t.pos() < blocks * s);
        self.tags = keep.into_iter().map(|t| Tag::new(t.pos() - blocks * s, t.key(), t.val().clone())).collect();
        input.consume(consume);
        o.produce(blocks * s, &emit);
        Ok(if wait_out != 0 { BlockRet::WaitForStream(&self.dst, wait_need) } else { BlockRet::WaitForStream(&self.src, wait_need) })
    }
}

/// GPU `RationalResampler<T>` for 4- and 8-byte samples (src/rational_resampler.rs:125-213).
pub struct CudaRationalResampler<T: rustradio::Sample> {
    h: *mut ffi::rrc_resampler_t,
    dev: i32,
    sin: Scratch,
    sout: Scratch,
    src: ReadStream<T>,
    dst: WriteStream<T>,
}
unsafe impl<T: rustradio::Sample> Send for CudaRationalResampler<T> {}

impl<T: rustradio::Sample> CudaRationalResampler<T> {
    pub fn new(src: ReadStream<T>, interp: usize, deci: usize) -> Result<(Self, ReadStream<T>)> {
        let mut h = ptr::null_mut();
        unsafe { check(ffi::rrc_resampler_create(0, std::mem::size_of::<T>(), interp, deci, &mut h))?; }   // Err on 0 (:130-135)
        let (dst, dr) = rustradio::stream::new_stream();
        Ok((Self { h, dev: 0, sin: Scratch::new(0), sout: Scratch::new(0), src, dst }, dr))
    }
}
impl<T: rustradio::Sample> Drop for CudaRationalResampler<T> { fn drop(&mut self) { unsafe { ffi::rrc_resampler_destroy(self.h); } } }
impl<T: rustradio::Sample> BlockName for CudaRationalResampler<T> { fn block_name(&self) -> &str { "CudaRationalResampler" } }
impl<T: rustradio::Sample> BlockEOF for CudaRationalResampler<T> {
    fn eof(&mut self) -> bool {
        let mut pending = 0;
        unsafe { ffi::rrc_resampler_state(self.h, ptr::null_mut(), ptr::null_mut(), ptr::null_mut(), &mut pending); }
        pending == 0 && self.src.eof()                       // src/rational_resampler.rs:209-213
    }
}
impl<T: rustradio::Sample> Block for CudaRationalResampler<T> {
    fn work(&mut self) -> Result<BlockRet<'_>> {
        let es = std::mem::size_of::<T>();
        let mut o = self.dst.write_buf()?;
        if o.is_empty() { return Ok(BlockRet::WaitForStream(&self.dst, 1)); }
        let (i, _tags) = self.src.read_buf()?;               // tags dropped like the reference (:156)
        let (din, dout) = (self.sin.reserve(i.len().max(1) * es)?, self.sout.reserve(o.len() * es)?);
        let (mut consumed, mut produced, mut wait_out) = (0usize, 0usize, 0i32);
        unsafe {
            check(ffi::rrc_memcpy_h2d(self.dev, din, i.slice().as_ptr().cast(), i.len() * es, ptr::null_mut()))?;
            check(ffi::rrc_resampler_run(self.h, din, i.len(), dout, o.len(), &mut consumed, &mut produced, &mut wait_out, ptr::null_mut()))?;
            check(ffi::rrc_memcpy_d2h(self.dev, o.slice().as_mut_ptr().cast(), dout, produced * es, ptr::null_mut()))?;
            check(ffi::rrc_stream_sync(self.dev, ptr::null_mut()))?;
        }
        i.consume(consumed);
        o.produce(produced, &[]);
        Ok(if wait_out != 0 { BlockRet::WaitForStream(&self.dst, 1) } else { BlockRet::WaitForStream(&self.src, 1) })
    }
}

/// GPU `QuadratureDemod`: `new(src, gain)` like the macro-generated src/quadrature_demod.rs:32-43.
pub struct CudaQuadratureDemod {
    gain: Float,
    dev: i32,
    sin: Scratch,
    sout: Scratch,
    src: ReadStream<Complex>,
    dst: WriteStream<Float>,
}
unsafe impl Send for CudaQuadratureDemod {}
impl CudaQuadratureDemod {
    pub fn new(src: ReadStream<Complex>, gain: Float) -> (Self, ReadStream<Float>) {
        let (dst, dr) = rustradio::stream::new_stream();
        (Self { gain, dev: 0, sin: Scratch::new(0), sout: Scratch::new(0), src, dst }, dr)
    }
}
impl BlockName for CudaQuadratureDemod { fn block_name(&self) -> &str { "CudaQuadratureDemod" } }
impl BlockEOF for CudaQuadratureDemod { fn eof(&mut self) -> bool { self.src.eof() } }
impl Block for CudaQuadratureDemod {
    fn work(&mut self) -> Result<BlockRet<'_>> {
        loop {
            let (inp, _) = self.src.read_buf()?;
            if inp.len() < 2 { return Ok(BlockRet::WaitForStream(&self.src, 2)); }
            let mut out = self.dst.write_buf()?;
            if out.is_empty() { return Ok(BlockRet::WaitForStream(&self.dst, 1)); }
            let n1 = (inp.len() - 1).min(out.len());
            let (din, dout) = (self.sin.reserve((n1 + 1) * 8)?, self.sout.reserve(n1 * 4)?);
            unsafe {
                check(ffi::rrc_memcpy_h2d(self.dev, din, inp.slice().as_ptr().cast(), (n1 + 1) * 8, ptr::null_mut()))?;
                check(ffi::rrc_quad_demod_run(self.dev, din.cast(), n1 + 1, self.gain, dout.cast(), ptr::null_mut()))?;
                check(ffi::rrc_memcpy_d2h(self.dev, out.slice().as_mut_ptr().cast(), dout, n1 * 4, ptr::null_mut()))?;
                check(ffi::rrc_stream_sync(self.dev, ptr::null_mut()))?;
            }
            inp.consume(n1);                                  // keeps one sample of history (:110)
            out.produce(n1, &[]);
        }
    }
}

/// GPU `RtlSdrDecode`: `new(src)` like the macro-generated src/rtlsdr_decode.rs:9-16.
pub struct CudaRtlSdrDecode {
    dev: i32,
    sin: Scratch,
    sout: Scratch,
    src: ReadStream<u8>,
    dst: WriteStream<Complex>,
}
unsafe impl Send for CudaRtlSdrDecode {}
impl CudaRtlSdrDecode {
    pub fn new(src: ReadStream<u8>) -> (Self, ReadStream<Complex>) {
        let (dst, dr) = rustradio::stream::new_stream();
        (Self { dev: 0, sin: Scratch::new(0), sout: Scratch::new(0), src, dst }, dr)
    }
}
impl BlockName for CudaRtlSdrDecode { fn block_name(&self) -> &str { "CudaRtlSdrDecode" } }
impl BlockEOF for CudaRtlSdrDecode { fn eof(&mut self) -> bool { self.src.eof() } }
impl Block for CudaRtlSdrDecode {
    fn work(&mut self) -> Result<BlockRet<'_>> {
        loop {
            let (inp, _) = self.src.read_buf()?;                  // tags dropped (:21)
            let isamples = inp.len() & !1;                        // :23
            if isamples == 0 { return Ok(BlockRet::WaitForStream(&self.src, 2)); }
            let mut out = self.dst.write_buf()?;
            if out.is_empty() { return Ok(BlockRet::WaitForStream(&self.dst, 1)); }
            let isamples = isamples.min(out.len() * 2);           // :32
            let osamples = isamples / 2;
            let (din, dout) = (self.sin.reserve(isamples)?, self.sout.reserve(osamples * 8)?);
            unsafe {
                check(ffi::rrc_memcpy_h2d(self.dev, din, inp.slice().as_ptr().cast(), isamples, ptr::null_mut()))?;
                check(ffi::rrc_rtlsdr_decode_run(self.dev, din.cast(), isamples, dout.cast(), ptr::null_mut()))?;
                check(ffi::rrc_memcpy_d2h(self.dev, out.slice().as_mut_ptr().cast(), dout, osamples * 8, ptr::null_mut()))?;
                check(ffi::rrc_stream_sync(self.dev, ptr::null_mut()))?;
            }
            inp.consume(isamples);
            out.produce(osamples, &[]);
        }
    }
}
