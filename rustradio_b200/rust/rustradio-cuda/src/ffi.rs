//! Raw bindings of include/rustradio_cuda.h (the kernel-level `rrc_*` entry points the blocks use).
//! UNCOMPILED in this repository's environment (no Rust toolchain); kept in sync with the header
//! by tests/test_abi.py::test_rust_ffi_declares_header_symbols.
#![allow(non_camel_case_types)]
use std::ffi::{c_char, c_float, c_int, c_uint, c_void};

#[repr(C)] pub struct rrc_fir_t { _p: [u8; 0] }
#[repr(C)] pub struct rrc_fftfilt_t { _p: [u8; 0] }
#[repr(C)] pub struct rrc_resampler_t { _p: [u8; 0] }
#[repr(C)] pub struct rrc_hilbert_t { _p: [u8; 0] }
#[repr(C)] pub struct rrc_iq_balance_t { _p: [u8; 0] }

pub const RRC_OK: c_int = 0;

unsafe extern "C" {
    pub fn rrc_abi_version() -> c_int;
    pub fn rrc_last_error() -> *const c_char;
    pub fn rrc_device_count(count: *mut c_int) -> c_int;
    pub fn rrc_malloc_device(device: c_int, bytes: usize, dev_ptr: *mut *mut c_void) -> c_int;
    pub fn rrc_free_device(device: c_int, dev_ptr: *mut c_void) -> c_int;
    pub fn rrc_malloc_pinned(bytes: usize, host_ptr: *mut *mut c_void) -> c_int;
    pub fn rrc_free_pinned(host_ptr: *mut c_void) -> c_int;
    pub fn rrc_memcpy_h2d(device: c_int, dev_dst: *mut c_void, host_src: *const c_void, bytes: usize, stream: *mut c_void) -> c_int;
    pub fn rrc_memcpy_d2h(device: c_int, host_dst: *mut c_void, dev_src: *const c_void, bytes: usize, stream: *mut c_void) -> c_int;
    pub fn rrc_memcpy_d2d(device: c_int, dev_dst: *mut c_void, dev_src: *const c_void, bytes: usize, stream: *mut c_void) -> c_int;
    pub fn rrc_stream_create(device: c_int, stream: *mut *mut c_void) -> c_int;
    pub fn rrc_stream_destroy(device: c_int, stream: *mut c_void) -> c_int;
    pub fn rrc_stream_sync(device: c_int, stream: *mut c_void) -> c_int;

    pub fn rrc_fir_c32_create(device: c_int, taps_c32: *const c_float, ntaps: usize, deci: usize, flags: c_uint, out: *mut *mut rrc_fir_t) -> c_int;
    pub fn rrc_fir_f32_create(device: c_int, taps: *const c_float, ntaps: usize, deci: usize, flags: c_uint, out: *mut *mut rrc_fir_t) -> c_int;
    pub fn rrc_fir_set_translate(h: *mut rrc_fir_t, samp_rate: c_float, freq: c_float) -> c_int;
    pub fn rrc_fir_destroy(h: *mut rrc_fir_t) -> c_int;
    pub fn rrc_fir_plan(ntaps: usize, deci: usize, in_len: usize, out_free: usize, consume: *mut usize, need: *mut usize,
                        out_n: *mut usize, wait_need: *mut usize, wait_on_output: *mut c_int) -> c_int;
    pub fn rrc_fir_run(h: *mut rrc_fir_t, in_dev: *const c_void, need: usize, out_dev: *mut c_void, out_n: usize, stream: *mut c_void) -> c_int;
    pub fn rrc_fir_c32_demod_run_batch(h: *mut rrc_fir_t, in_dev: *const c_void, in_stride: usize, need: usize, gain: c_float,
                                       out_dev: *mut c_float, out_stride: usize, out_n: usize, nchan: usize, stream: *mut c_void) -> c_int;

    pub fn rrc_fftfilt_c32_create(device: c_int, taps_c32: *const c_float, ntaps: usize, out: *mut *mut rrc_fftfilt_t) -> c_int;
    pub fn rrc_fftfilt_destroy(h: *mut rrc_fftfilt_t) -> c_int;
    pub fn rrc_fftfilt_ref_fft_size(ntaps: usize, fft_size: *mut usize, nsamples: *mut usize) -> c_int;
    pub fn rrc_fftfilt_plan(ntaps: usize, buffered: usize, in_len: usize, out_free: usize, blocks: *mut usize, consume: *mut usize,
                            buffered_after: *mut usize, wait_need: *mut usize, wait_on_output: *mut c_int) -> c_int;
    pub fn rrc_fftfilt_set_history(h: *mut rrc_fftfilt_t, hist_dev_c32: *const c_float, n_samples: usize, stream: *mut c_void) -> c_int;
    pub fn rrc_fftfilt_run(h: *mut rrc_fftfilt_t, in_dev: *const c_float, n: usize, out_dev: *mut c_float, stream: *mut c_void) -> c_int;

    pub fn rrc_resampler_create(device: c_int, elem_size: usize, interp: usize, deci: usize, out: *mut *mut rrc_resampler_t) -> c_int;
    pub fn rrc_resampler_destroy(h: *mut rrc_resampler_t) -> c_int;
    pub fn rrc_resampler_state(h: *const rrc_resampler_t, interp: *mut i64, deci: *mut i64, counter: *mut i64, has_pending: *mut c_int) -> c_int;
    pub fn rrc_resampler_run(h: *mut rrc_resampler_t, in_dev: *const c_void, n_in: usize, out_dev: *mut c_void, out_cap: usize,
                             consumed: *mut usize, produced: *mut usize, wait_on_output: *mut c_int, stream: *mut c_void) -> c_int;

    pub fn rrc_quad_demod_run(device: c_int, in_dev_c32: *const c_float, n_in: usize, gain: c_float, out_dev: *mut c_float, stream: *mut c_void) -> c_int;

    // RtlSdrDecode::work                        src/rtlsdr_decode.rs:18-48
    pub fn rrc_rtlsdr_decode_plan(in_len_bytes: usize, out_free: usize, consume_bytes: *mut usize, produce: *mut usize,
                                  wait_need: *mut usize, wait_on_output: *mut c_int) -> c_int;
    pub fn rrc_rtlsdr_decode_run(device: c_int, in_dev: *const u8, n_bytes: usize, out_dev_c32: *mut c_float, stream: *mut c_void) -> c_int;
    // RtlSdrDecode fused into the first load of the filters (u8 I/Q input mode)
    pub fn rrc_fir_set_input_u8iq(h: *mut rrc_fir_t, on: c_int) -> c_int;
    pub fn rrc_fir_uses_tensor_cores(h: *const rrc_fir_t, yes: *mut c_int) -> c_int;
    pub fn rrc_fftfilt_set_input_u8iq(h: *mut rrc_fftfilt_t, on: c_int) -> c_int;
    // WindowType::make_window / fir::hilbert / Hilbert::work     src/window.rs:63-185, src/fir.rs:660-680, src/hilbert.rs:72-128
    pub fn rrc_make_window(window_type: c_int, parm: c_float, ntaps: usize, window_out: *mut c_float) -> c_int;
    pub fn rrc_hilbert_taps(window: *const c_float, ntaps: usize, taps_out: *mut c_float) -> c_int;
    pub fn rrc_hilbert_create(device: c_int, taps: *const c_float, ntaps: usize, out: *mut *mut rrc_hilbert_t) -> c_int;
    pub fn rrc_hilbert_destroy(h: *mut rrc_hilbert_t) -> c_int;
    pub fn rrc_hilbert_run(h: *mut rrc_hilbert_t, in_dev: *const c_float, n: usize, out_dev_c32: *mut c_float, stream: *mut c_void) -> c_int;
    // the `sync` blocks next to the filters      src/multiply_const.rs, add_const.rs, complex_to_mag2.rs, tee.rs, iq_balance.rs
    pub fn rrc_multiply_const_f32_run(device: c_int, in_dev: *const c_float, n: usize, val: c_float, out_dev: *mut c_float, stream: *mut c_void) -> c_int;
    pub fn rrc_multiply_const_c32_run(device: c_int, in_dev_c32: *const c_float, n: usize, val_re: c_float, val_im: c_float,
                                      out_dev_c32: *mut c_float, stream: *mut c_void) -> c_int;
    pub fn rrc_add_const_f32_run(device: c_int, in_dev: *const c_float, n: usize, val: c_float, out_dev: *mut c_float, stream: *mut c_void) -> c_int;
    pub fn rrc_add_const_c32_run(device: c_int, in_dev_c32: *const c_float, n: usize, val_re: c_float, val_im: c_float,
                                 out_dev_c32: *mut c_float, stream: *mut c_void) -> c_int;
    pub fn rrc_complex_to_mag2_run(device: c_int, in_dev_c32: *const c_float, n: usize, out_dev: *mut c_float, stream: *mut c_void) -> c_int;
    pub fn rrc_tee_run(device: c_int, in_dev: *const c_void, nbytes: usize, out1_dev: *mut c_void, out2_dev: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn rrc_iq_balance_alpha_from_tau(sample_rate: c_uint, tau_seconds: f64, alpha: *mut c_float) -> c_int;
    pub fn rrc_iq_balance_create(device: c_int, alpha: c_float, out: *mut *mut rrc_iq_balance_t) -> c_int;
    pub fn rrc_iq_balance_destroy(h: *mut rrc_iq_balance_t) -> c_int;
    pub fn rrc_iq_balance_run(h: *mut rrc_iq_balance_t, in_dev_c32: *const c_float, n: usize, out_dev_c32: *mut c_float, stream: *mut c_void) -> c_int;
}

/// Turn a status code into rustradio's `Error::DeviceError`-style error (src/lib.rs:288-294).
pub fn check(code: c_int) -> rustradio::Result<()> {
    if code == RRC_OK {
        return Ok(());
    }
    // SAFETY: rrc_last_error returns a NUL-terminated thread-local buffer.
    let msg = unsafe { std::ffi::CStr::from_ptr(rrc_last_error()) }.to_string_lossy().into_owned();
    Err(rustradio::Error::msg(format!("rustradio-cuda error {code}: {msg}")))
}
