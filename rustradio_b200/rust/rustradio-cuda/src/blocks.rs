//! The GPU blocks behind rustradio's `Block` trait (src/block.rs:115-126): `CudaFirFilter<T>` with its builder
//! (`deci`, `translate`), `CudaFftFilter` / `CudaFftFilterFloat`, `CudaRationalResampler<T>` with its typestate
//! builder, `CudaQuadratureDemod`, `CudaRtlSdrDecode`, `CudaRtlSdrEncode`.
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

/// Sample types the GPU FIR / FFT filters exist for: `Complex` (`rrc_*_c32_create`) and `Float`
/// (`rrc_*_f32_create`) — `FirFilter<Complex>`, `FirFilter<Float>`, `FftFilter`, `FftFilterFloat`.
pub trait GpuSample: rustradio::Sample<Type = Self> + Copy + Send + 'static {
    const NAME: &'static str;
    /// # Safety: plain FFI.
    unsafe fn fir_create(dev: i32, taps: &[Self], deci: usize, flags: u32, out: *mut *mut ffi::rrc_fir_t) -> i32;
    unsafe fn fftfilt_create(dev: i32, taps: &[Self], out: *mut *mut ffi::rrc_fftfilt_t) -> i32;
}
impl GpuSample for Complex {
    const NAME: &'static str = "Complex";
    unsafe fn fir_create(dev: i32, taps: &[Self], deci: usize, flags: u32, out: *mut *mut ffi::rrc_fir_t) -> i32 {
        // SAFETY: Complex<f32> is repr(C) (re, im).
        unsafe { ffi::rrc_fir_c32_create(dev, taps.as_ptr().cast(), taps.len(), deci, flags, out) }
    }
    unsafe fn fftfilt_create(dev: i32, taps: &[Self], out: *mut *mut ffi::rrc_fftfilt_t) -> i32 {
        unsafe { ffi::rrc_fftfilt_c32_create(dev, taps.as_ptr().cast(), taps.len(), out) }
    }
}
impl GpuSample for Float {
    const NAME: &'static str = "Float";
    unsafe fn fir_create(dev: i32, taps: &[Self], deci: usize, flags: u32, out: *mut *mut ffi::rrc_fir_t) -> i32 {
        unsafe { ffi::rrc_fir_f32_create(dev, taps.as_ptr(), taps.len(), deci, flags, out) }
    }
    unsafe fn fftfilt_create(dev: i32, taps: &[Self], out: *mut *mut ffi::rrc_fftfilt_t) -> i32 {
        unsafe { ffi::rrc_fftfilt_f32_create(dev, taps.as_ptr(), taps.len(), out) }
    }
}

/// Builder for the GPU FIR filter block: `FirFilterBuilder<T>` (src/fir.rs:303-340) — `deci()`, and on
/// `Complex` `translate()` (src/fir.rs:476-486) — plus `device()` / `flags()` for the GPU side.
pub struct CudaFirFilterBuilder<T: GpuSample> {
    taps: Vec<T>,
    deci: usize,
    translate: Option<(Float, Float)>,
    dev: i32,
    flags: u32,
}
impl<T: GpuSample> CudaFirFilterBuilder<T> {
    /// Set the decimation (default 1).  Panics on 0 like src/fir.rs:319.
    #[must_use]
    pub fn deci(mut self, deci: usize) -> Self {
        assert_ne!(deci, 0);
        self.deci = deci;
        self
    }
    /// CUDA device ordinal (default 0).
    #[must_use]
    pub fn device(mut self, dev: i32) -> Self {
        self.dev = dev;
        self
    }
    /// `RRC_FIR_*` flags (e.g. `RRC_FIR_NO_TENSOR` to keep the FP32 kernels).
    #[must_use]
    pub fn flags(mut self, flags: u32) -> Self {
        self.flags = flags;
        self
    }
    /// Build the block (src/fir.rs:323-339).
    #[must_use]
    pub fn build(self, src: ReadStream<T>) -> (CudaFirFilter<T>, ReadStream<T>) {
        assert!(!self.taps.is_empty());                     // src/fir.rs:372
        let mut h = ptr::null_mut();
        unsafe {
            check(T::fir_create(self.dev, &self.taps, self.deci, self.flags, &mut h)).expect("rrc_fir_create");
            if let Some((samp_rate, freq)) = self.translate {
                assert!(samp_rate > 0.0);                   // src/fir.rs:436
                // same tap pre-rotation and per-output rotator as T::new_translator (src/fir.rs:427-462)
                check(ffi::rrc_fir_set_translate(h, samp_rate, freq)).expect("rrc_fir_set_translate");
            }
        }
        let (dst, dr) = rustradio::stream::new_stream();
        let es = std::mem::size_of::<T>();
        (CudaFirFilter { h, ntaps: self.taps.len(), deci: self.deci, es, dev: self.dev, sin: Scratch::new(self.dev),
                         sout: Scratch::new(self.dev), src, dst }, dr)
    }
}
impl CudaFirFilterBuilder<Complex> {
    /// Mix by `-freq` Hz while filtering (src/fir.rs:476-486).
    #[must_use]
    pub fn translate(mut self, samp_rate: Float, freq: Float) -> Self {
        self.translate = Some((samp_rate, freq));
        self
    }
}

/// GPU `FirFilter<T>` for `T = Complex` and `T = Float`: `new(src, taps)` / `builder(taps)` as
/// src/fir.rs:357-386.
pub struct CudaFirFilter<T: GpuSample> {
    h: *mut ffi::rrc_fir_t,
    ntaps: usize,
    deci: usize,
    es: usize,
    dev: i32,
    sin: Scratch,
    sout: Scratch,
    src: ReadStream<T>,
    dst: WriteStream<T>,
}
unsafe impl<T: GpuSample> Send for CudaFirFilter<T> {}

impl<T: GpuSample> CudaFirFilter<T> {
    /// `FirFilter::builder(taps)` (src/fir.rs:362-368).
    pub fn builder(taps: impl Into<Vec<T>>) -> CudaFirFilterBuilder<T> {
        CudaFirFilterBuilder { taps: taps.into(), deci: 1, translate: None, dev: 0, flags: 0 }
    }
    /// `FirFilter::new(src, taps)` (src/fir.rs:370-386).
    pub fn new(src: ReadStream<T>, taps: impl AsRef<[T]>) -> (Self, ReadStream<T>) {
        Self::builder(taps.as_ref().to_vec()).build(src)
    }
    pub fn with_deci(src: ReadStream<T>, taps: impl AsRef<[T]>, deci: usize) -> (Self, ReadStream<T>) {
        Self::builder(taps.as_ref().to_vec()).deci(deci).build(src)
    }
}
impl<T: GpuSample> Drop for CudaFirFilter<T> { fn drop(&mut self) { unsafe { ffi::rrc_fir_destroy(self.h); } } }
impl<T: GpuSample> BlockName for CudaFirFilter<T> {
    fn block_name(&self) -> &str { if T::NAME == "Complex" { "CudaFirFilter<Complex>" } else { "CudaFirFilter<Float>" } }
}
impl<T: GpuSample> BlockEOF for CudaFirFilter<T> { fn eof(&mut self) -> bool { self.src.eof() } }

impl<T: GpuSample> Block for CudaFirFilter<T> {
    fn work(&mut self) -> Result<BlockRet<'_>> {
        let (input, mut tags) = self.src.read_buf()?;
        let mut out = self.dst.write_buf()?;
        let (mut n, mut need, mut out_n, mut wait_need, mut wait_out) = (0usize, 0usize, 0usize, 0usize, 0i32);
        // The integer part of FirFilter::work (src/fir.rs:496-525), shared with the C++ mirror.
        unsafe { check(ffi::rrc_fir_plan(self.ntaps, self.deci, input.len(), out.len(), &mut n, &mut need, &mut out_n, &mut wait_need, &mut wait_out))?; }
        if n == 0 {
            return Ok(if wait_out != 0 { BlockRet::WaitForStream(&self.dst, wait_need) } else { BlockRet::WaitForStream(&self.src, wait_need) });
        }
        let es = self.es;
        let (din, dout) = (self.sin.reserve(need * es)?, self.sout.reserve(out_n * es)?);
        unsafe {
            check(ffi::rrc_memcpy_h2d(self.dev, din, input.slice().as_ptr().cast(), need * es, ptr::null_mut()))?;
            check(ffi::rrc_fir_run(self.h, din, need, dout, out_n, ptr::null_mut()))?;
            check(ffi::rrc_memcpy_d2h(self.dev, out.slice().as_mut_ptr().cast(), dout, out_n * es, ptr::null_mut()))?;
            check(ffi::rrc_stream_sync(self.dev, ptr::null_mut()))?;
        }
        tags.retain(|t| t.pos() < n);                       // src/fir.rs:536
        input.consume(n);
        if self.deci != 1 { for t in &mut tags { t.set_pos(t.pos() / self.deci); } }
        out.produce(out_n, &tags);
        Ok(BlockRet::Again)
    }
}

/// GPU `FftFilter` (`T = Complex`, src/fft_filter.rs:241-255) and `FftFilterFloat` (`T = Float`,
/// src/fft_filter.rs:365-426): `new(src, taps)`; same count rule (whole blocks of `nsamples`, partial block
/// retained) via `rrc_fftfilt_plan`.  The Float form runs the device's real-stream mode (two real blocks per
/// complex transform) instead of the reference's widen -> complex filter -> `.re`; counts, tags and BlockRets
/// per `work()` are those of the inner complex filter on the same number of samples.
pub struct CudaFftFilterT<T: GpuSample> {
    h: *mut ffi::rrc_fftfilt_t,
    ntaps: usize,
    nsamples: usize,
    buf: Vec<T>,                        // self.buf of the reference (host side, < nsamples samples)
    tags: Vec<rustradio::stream::Tag>,
    dev: i32,
    sin: Scratch,
    sout: Scratch,
    es: usize,
    src: ReadStream<T>,
    dst: WriteStream<T>,
}
unsafe impl<T: GpuSample> Send for CudaFftFilterT<T> {}
/// `FftFilter` on the GPU.
pub type CudaFftFilter = CudaFftFilterT<Complex>;
/// `FftFilterFloat` on the GPU (`new(src: ReadStream<Float>, taps: &[Float])`).
pub type CudaFftFilterFloat = CudaFftFilterT<Float>;

impl<T: GpuSample> CudaFftFilterT<T> {
    pub fn new<V: Into<Vec<T>>>(src: ReadStream<T>, taps: V) -> (Self, ReadStream<T>) {
        let taps: Vec<T> = taps.into();
        assert!(!taps.is_empty());
        let mut h = ptr::null_mut();
        let (mut fft_size, mut nsamples) = (0usize, 0usize);
        unsafe {
            check(T::fftfilt_create(0, &taps, &mut h)).expect("rrc_fftfilt_create");
            check(ffi::rrc_fftfilt_ref_fft_size(taps.len(), &mut fft_size, &mut nsamples)).unwrap();
        }
        let (dst, dr) = rustradio::stream::new_stream();
        (Self { h, ntaps: taps.len(), nsamples, buf: Vec::with_capacity(nsamples), tags: Vec::new(), dev: 0,
                sin: Scratch::new(0), sout: Scratch::new(0), es: std::mem::size_of::<T>(), src, dst }, dr)
    }
}
impl<T: GpuSample> Drop for CudaFftFilterT<T> { fn drop(&mut self) { unsafe { ffi::rrc_fftfilt_destroy(self.h); } } }
impl<T: GpuSample> BlockName for CudaFftFilterT<T> {
    fn block_name(&self) -> &str { if T::NAME == "Complex" { "CudaFftFilter" } else { "CudaFftFilterFloat" } }
}
impl<T: GpuSample> BlockEOF for CudaFftFilterT<T> { fn eof(&mut self) -> bool { self.src.eof() } }

impl<T: GpuSample> Block for CudaFftFilterT<T> {
    fn work(&mut self) -> Result<BlockRet<'_>> {
        use rustradio::stream::Tag;
        let mut o = self.dst.write_buf()?;
        let (input, tags) = self.src.read_buf()?;
        let (mut blocks, mut consume, mut after, mut wait_need, mut wait_out) = (0usize, 0usize, 0usize, 0usize, 0i32);
        unsafe { check(ffi::rrc_fftfilt_plan(self.ntaps, self.buf.len(), input.len(), o.len(), &mut blocks, &mut consume, &mut after, &mut wait_need, &mut wait_out))?; }
        let s = self.nsamples;
        let buffered = self.buf.len();
        // samples handed to the filter: (buf ++ input[..consume]); the first blocks*s of them are filtered now
        let mut staged: Vec<T> = Vec::with_capacity(buffered + consume);
        staged.extend_from_slice(&self.buf);
        staged.extend_from_slice(&input.slice()[..consume]);
        if blocks > 0 {
            let n = blocks * s;
            let es = self.es;
            let (din, dout) = (self.sin.reserve(n * es)?, self.sout.reserve(n * es)?);
            unsafe {
                check(ffi::rrc_memcpy_h2d(self.dev, din, staged.as_ptr().cast(), n * es, ptr::null_mut()))?;
                check(ffi::rrc_fftfilt_run(self.h, din.cast(), n, dout.cast(), ptr::null_mut()))?;
                check(ffi::rrc_memcpy_d2h(self.dev, o.slice().as_mut_ptr().cast(), dout, n * es, ptr::null_mut()))?;
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

/// Typestate builder of the resampler, `RationalResamplerBuilder<T>` (src/rational_resampler.rs:19-92):
/// `CudaRationalResampler::<T>::builder().interp(i).deci(d).build(src)` (either order).
pub struct CudaRationalResamplerBuilder<T> { dummy: std::marker::PhantomData<T> }
pub struct CudaRationalResamplerBuilderInterp<T> { dummy: std::marker::PhantomData<T>, interp: usize }
pub struct CudaRationalResamplerBuilderDeci<T> { dummy: std::marker::PhantomData<T>, deci: usize }
pub struct CudaRationalResamplerBuilderBoth<T> { dummy: std::marker::PhantomData<T>, interp: usize, deci: usize }
impl<T> Default for CudaRationalResamplerBuilder<T> { fn default() -> Self { Self::new() } }
impl<T> CudaRationalResamplerBuilder<T> {
    #[must_use]
    pub fn new() -> Self { Self { dummy: std::marker::PhantomData } }
    #[must_use]
    pub fn deci(self, deci: usize) -> CudaRationalResamplerBuilderDeci<T> { CudaRationalResamplerBuilderDeci { deci, dummy: self.dummy } }
    #[must_use]
    pub fn interp(self, interp: usize) -> CudaRationalResamplerBuilderInterp<T> { CudaRationalResamplerBuilderInterp { interp, dummy: self.dummy } }
}
impl<T> CudaRationalResamplerBuilderInterp<T> {
    #[must_use]
    pub fn deci(self, deci: usize) -> CudaRationalResamplerBuilderBoth<T> { CudaRationalResamplerBuilderBoth { interp: self.interp, deci, dummy: self.dummy } }
}
impl<T> CudaRationalResamplerBuilderDeci<T> {
    #[must_use]
    pub fn interp(self, interp: usize) -> CudaRationalResamplerBuilderBoth<T> { CudaRationalResamplerBuilderBoth { deci: self.deci, interp, dummy: self.dummy } }
}
impl<T: rustradio::Sample> CudaRationalResamplerBuilderBoth<T> {
    /// Errors if the interpolation or decimation value is 0 (src/rational_resampler.rs:85-91).
    pub fn build(self, src: ReadStream<T>) -> Result<(CudaRationalResampler<T>, ReadStream<T>)> {
        CudaRationalResampler::new(src, self.interp, self.deci)
    }
}

/// GPU `RationalResampler<T>` for 1/2/4/8/16-byte samples (src/rational_resampler.rs:125-213).  The carried
/// `counter` / `pending` state lives in the handle; `eof()` also requires `pending.is_none()` (:209-213).
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
    /// `RationalResampler::builder()` (src/rational_resampler.rs:113-117).
    #[must_use]
    pub fn builder() -> CudaRationalResamplerBuilder<T> { CudaRationalResamplerBuilder::<T>::new() }
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

/// GPU `RtlSdrEncode`: `new(src)` like the macro-generated src/rtlsdr_encode.rs:12-20.
pub struct CudaRtlSdrEncode {
    dev: i32,
    sin: Scratch,
    sout: Scratch,
    src: ReadStream<Complex>,
    dst: WriteStream<u8>,
}
unsafe impl Send for CudaRtlSdrEncode {}
impl CudaRtlSdrEncode {
    pub fn new(src: ReadStream<Complex>) -> (Self, ReadStream<u8>) {
        let (dst, dr) = rustradio::stream::new_stream();
        (Self { dev: 0, sin: Scratch::new(0), sout: Scratch::new(0), src, dst }, dr)
    }
}
impl BlockName for CudaRtlSdrEncode { fn block_name(&self) -> &str { "CudaRtlSdrEncode" } }
impl BlockEOF for CudaRtlSdrEncode { fn eof(&mut self) -> bool { self.src.eof() } }
impl Block for CudaRtlSdrEncode {
    fn work(&mut self) -> Result<BlockRet<'_>> {
        loop {
            let (inp, _) = self.src.read_buf()?;                  // tags dropped (:31)
            if inp.is_empty() { return Ok(BlockRet::WaitForStream(&self.src, 1)); }
            let mut out = self.dst.write_buf()?;
            if out.len() < 2 { return Ok(BlockRet::WaitForStream(&self.dst, 2)); }
            let isamples = inp.len().min(out.len() / 2);          // :41
            let obytes = isamples * 2;
            let (din, dout) = (self.sin.reserve(isamples * 8)?, self.sout.reserve(obytes)?);
            unsafe {
                check(ffi::rrc_memcpy_h2d(self.dev, din, inp.slice().as_ptr().cast(), isamples * 8, ptr::null_mut()))?;
                check(ffi::rrc_rtlsdr_encode_run(self.dev, din.cast(), isamples, dout.cast(), ptr::null_mut()))?;
                check(ffi::rrc_memcpy_d2h(self.dev, out.slice().as_mut_ptr().cast(), dout, obytes, ptr::null_mut()))?;
                check(ffi::rrc_stream_sync(self.dev, ptr::null_mut()))?;
            }
            inp.consume(isamples);
            out.produce(obytes, &[]);
        }
    }
}
