/*
 * rustradio_cuda.h — C ABI of the B200 filtering hot path for rustradio.
 *
 * This is the whole drop-in boundary: `extern "C"`, plain pointers and
 * sizes, no C++/torch/Rust types.  rustradio has no FFI today (it is pure
 * Rust); these are the entry points the `rustradio-cuda` crate's FFI binds
 * (see INTEGRATION.md and rustradio_b200/rust/rustradio-cuda/src/ffi.rs).
 * Each group cites the reference code it replaces (paths relative to the
 * rustradio v0.18.2 source tree).
 *
 * Conventions
 *   - Every function returns an `int` status: RRC_OK (0) or a negative
 *     RRC_ERR_*; the text of the last error on the calling thread is
 *     rrc_last_error().  Nothing unwinds across the boundary.
 *   - Complex<f32> samples/taps are interleaved (re, im) floats, i.e.
 *     num_complex::Complex<f32>'s #[repr(C)] layout; `float*` arguments that
 *     carry c32 data have 2 floats per sample and sizes count SAMPLES.
 *   - "dev" pointers are device pointers on the handle's device; "host"
 *     pointers are host memory (pinned memory makes the copies asynchronous).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the handle's device
 *     default stream).  Kernel-level *_run calls only enqueue; they do not
 *     synchronise.  *_run_host calls return after the result is in host memory.
 *   - Every entry point selects the handle's device itself (rustradio's
 *     AsyncGraph may call work() from different threads, src/agraph.rs:64-97).
 *   - There is no CPU fallback anywhere behind this ABI: without a CUDA
 *     device every compute entry point fails with RRC_ERR_CUDA.
 */
#ifndef RUSTRADIO_CUDA_H
#define RUSTRADIO_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RRC_OK            0
#define RRC_ERR_INVALID  (-1)   /* bad argument (the reference would assert!/Err) */
#define RRC_ERR_CUDA     (-2)   /* CUDA runtime/driver failure -> Error::DeviceError (src/lib.rs:288-294) */
#define RRC_ERR_NOMEM    (-3)
#define RRC_ERR_STATE    (-4)   /* call made in the wrong state */
#define RRC_ERR_UNSUPPORTED (-5)

#define RRC_ABI_VERSION 1

/* ------------------------------------------------------------ runtime --- */
int         rrc_abi_version(void);
const char* rrc_last_error(void);
int rrc_device_count(int* count);
int rrc_device_name(int device, char* buf, size_t buflen);
int rrc_device_sm_count(int device, int* sms);

/* Device / pinned memory and copies, so the host language needs no cudart. */
int rrc_malloc_device(int device, size_t bytes, void** dev_ptr);
int rrc_free_device(int device, void* dev_ptr);
int rrc_malloc_pinned(size_t bytes, void** host_ptr);
int rrc_free_pinned(void* host_ptr);
/* Pinned host memory placed on the NUMA node of `device`'s PCIe root complex (sysfs numa_node of its bus id;
 * anonymous mapping + mbind(MPOL_PREFERRED) + cudaHostRegister); plain rrc_malloc_pinned when the platform
 * exposes no NUMA information.  Free with rrc_free_pinned. */
int rrc_malloc_pinned_near(int device, size_t bytes, void** host_ptr);
int rrc_device_numa_node(int device, int* node);             /* -1: unknown */
int rrc_host_register(void* host_ptr, size_t bytes);      /* pin caller memory */
int rrc_host_unregister(void* host_ptr);
int rrc_memset_device(int device, void* dev_ptr, int value, size_t bytes, void* stream);
int rrc_memcpy_h2d(int device, void* dev_dst, const void* host_src, size_t bytes, void* stream);
int rrc_memcpy_d2h(int device, void* host_dst, const void* dev_src, size_t bytes, void* stream);
int rrc_memcpy_d2d(int device, void* dev_dst, const void* dev_src, size_t bytes, void* stream);
/* Peer access for time-segment shards: halos travel GPU to GPU over NVLink with plain loads /
 * copies, no collective.  One process driving several GPUs: rrc_peer_enable + any device pointer of `peer`
 * is then valid in kernels and copies on `device`.  One process per GPU: export a cudaMalloc'ed buffer
 * (rrc_malloc_device) as a 64-byte IPC handle, ship it to the neighbour by any means, open it there. */
int rrc_peer_enable(int device, int peer);
int rrc_memcpy_peer(int dst_device, void* dst, int src_device, const void* src, size_t bytes, void* stream);
int rrc_ipc_export(int device, void* dev_ptr, unsigned char handle[64]);
int rrc_ipc_open(int device, const unsigned char handle[64], void** dev_ptr);
int rrc_ipc_close(int device, void* dev_ptr);
int rrc_stream_create(int device, void** stream);
int rrc_stream_destroy(int device, void* stream);
int rrc_stream_sync(int device, void* stream);
int rrc_device_sync(int device);
/* CUDA-event timing on a stream (used by bench.py; torch.cuda.Event only sees torch's stream). */
int rrc_event_create(int device, void** event);
int rrc_event_destroy(int device, void* event);
int rrc_event_record(int device, void* event, void* stream);
/* Stream-ordered visibility between streams: work enqueued on `stream` after this call runs
 * after everything recorded into `event`. */
int rrc_event_wait(int device, void* event, void* stream);
int rrc_event_sync(int device, void* event);
int rrc_event_elapsed_ms(int device, void* start, void* stop, float* ms);
/* Fill a device buffer with the deterministic synthetic signal used by the
 * tests and bench (same splitmix64 counter generator as oracle/rr_oracle.c:
 * float index i -> U(-1,1)); n_floats floats starting at float index first. */
int rrc_synth_f32(int device, uint64_t seed, uint64_t first_index, float* dev_out, size_t n_floats, void* stream);
/* Number of kernels this library has launched on the calling process so far. */
int rrc_launch_count(uint64_t* launches);

/* ---------------------------------------------------------------- FIR --- */
/*
 * Replaces Fir<T>::new / filter / filter_n_inplace (src/fir.rs:156-197) and the
 * compute step of FirFilter<T>::work (src/fir.rs:526-531) for T = Complex and
 * T = Float.  out[i] = sum_{j<ntaps} in[i*deci + j] * taps[ntaps-1-j].
 * The host computes n/need/out_n exactly as src/fir.rs:496-525 (see
 * rrc_fir_plan) and calls run with need = (out_n-1)*deci + ntaps valid inputs.
 */
typedef struct rrc_fir rrc_fir_t;

/* flags */
#define RRC_FIR_NO_REAL_TAP_FASTPATH 1u  /* always do the full complex*complex MAC even if every tap has im == 0 */
#define RRC_FIR_FORCE_GENERIC        2u  /* use the one-thread-per-output fallback kernel (testing) */
#define RRC_FIR_NO_TENSOR            4u  /* keep the FP32 kernels: no fp16x3 tensor-core Toeplitz product */

int rrc_fir_c32_create(int device, const float* taps_c32, size_t ntaps, size_t deci, unsigned flags, rrc_fir_t** out);
int rrc_fir_f32_create(int device, const float* taps, size_t ntaps, size_t deci, unsigned flags, rrc_fir_t** out);
/* FirFilterBuilder::translate (src/fir.rs:476-486) + new_translator (:427-462):
 * rotates the taps (same f32 recurrence) and arms the per-output rotator.
 * The rotator phase is evaluated per output from the exact f64 angle instead
 * of the reference's drifting f32 recurrence (:464-473).
 * freq == 0 is a no-op like the reference (:438-440).  c32 only. */
int rrc_fir_set_translate(rrc_fir_t* h, float samp_rate, float freq);
int rrc_fir_destroy(rrc_fir_t* h);
int rrc_fir_ntaps(const rrc_fir_t* h, size_t* ntaps);
int rrc_fir_deci(const rrc_fir_t* h, size_t* deci);
/* 1 if the real-tap fast path (2 FMA per tap instead of 4) is active. */
int rrc_fir_uses_real_taps(const rrc_fir_t* h, int* yes);
/* 1 if runs go through a tensor-core Toeplitz kernel: at least 16 taps and
 *  - deci 1, 2, 4 or 8 with 7*deci + ntaps <= 320 (the "walk" kernels): c32 samples with real taps, c32 samples with
 *    complex taps (translate filters included) — both also from u8 I/Q input —, or f32 streams; or
 *  - c32 samples, real taps, ntaps >= 32 * deci, no translate (generic kernel; other decimations / longer filters).
 * Samples (per warp tile) and taps are scaled by powers of two and split hi + lo in fp16 (22 significant bits);
 * every product is hi*hi + hi*lo + lo*hi with FP32 accumulation.  FP32-class accuracy: rel-RMS error 1e-7..1e-6
 * against the f64 convolution, next to 1e-7..3e-7 for the sequential f32 loop (bar 1e-5).  Declared like the
 * real-tap fast path; RRC_FIR_NO_TENSOR (flag) or RRC_FIR_TENSOR=0 (environment) disables it. */
int rrc_fir_uses_tensor_cores(const rrc_fir_t* h, int* yes);
/* Store epilogues: the sample-wise neighbour that follows a Complex filter in the graph is
 * applied by the filter's own output store instead of costing a block and a full HBM round trip:
 *   RRC_EPI_MULTIPLY_CONST  y * (re + i im)   MultiplyConst<Complex>   src/multiply_const.rs:16-23
 *   RRC_EPI_ADD_CONST       y + (re + i im)   AddConst<Complex>        src/add_const.rs:36-44
 *   RRC_EPI_MAG2            |y|^2             ComplexToMag2            src/complex_to_mag2.rs:17-20  (out becomes f32)
 * Same separately rounded operations as the stand-alone kernels: fused == the two blocks back to back, bit for
 * bit.  (QuadratureDemod -> MultiplyConst<Float> is the `gain` of rrc_fir_c32_demod_run_batch.)  A FIR handle
 * with an epilogue runs on the FP32 kernels (rrc_fir_uses_tensor_cores reports 0). */
#define RRC_EPI_NONE           0
#define RRC_EPI_MULTIPLY_CONST 1
#define RRC_EPI_ADD_CONST      2
#define RRC_EPI_MAG2           3
int rrc_fir_set_epilogue(rrc_fir_t* h, int kind, float re, float im);
/* Which kernel the planner chose for this handle (reports / INTEGRATION.md): e.g. "fir_tc1_kernel<KS=5,D=1> ...",
 * "fir_rtu_kernel<D=10,QB=26,R=8> ..." (real-tap decimating filters with deci 5 or 10 and <= 26 taps per polyphase
 * branch: taps travel as kernel parameters and reach FFMA2 as uniform-register operands), "fir_rt_kernel<...>".
 * c32 filters with real taps, deci 1 and <= 65 taps have two kernels and pick per launch: "fir_tc5_kernel<KS=..>"
 * (tcgen05.mma, taps and accumulators in tensor memory) for launches of at least 3 tiles of 8192 outputs per SM,
 * "fir_tc1_kernel" (mma.sync) below that and for the fused demod; after a launch the name is the one that ran.
 * RRC_FIR_TCGEN05=0 (environment) keeps such filters on mma.sync. */
int rrc_fir_kernel_name(const rrc_fir_t* h, char* buf, size_t buflen);
/* Restart the translate rotator's output counter (new stream). */
int rrc_fir_reset(rrc_fir_t* h);

/* The integer part of FirFilter::work (src/fir.rs:496-525): given the input
 * window length and the free output space, how many samples to consume, how
 * many inputs the kernel must see and how many outputs it produces.
 * consume == 0 means WaitForStream: *wait_need is the sample count to wait
 * for and *wait_on_output says which stream (0 = src, 1 = dst). */
int rrc_fir_plan(size_t ntaps, size_t deci, size_t in_len, size_t out_free,
                 size_t* consume, size_t* need, size_t* out_n, size_t* wait_need, int* wait_on_output);

int rrc_fir_run(rrc_fir_t* h, const void* in_dev, size_t need, void* out_dev, size_t out_n, void* stream);
/* nchan independent channels with the same taps: channel c reads
 * in_dev + c*in_stride samples and writes out_dev + c*out_stride samples. */
int rrc_fir_run_batch(rrc_fir_t* h, const void* in_dev, size_t in_stride, size_t need,
                      void* out_dev, size_t out_stride, size_t out_n, size_t nchan, void* stream);
/* Fused FirFilter<Complex> -> QuadratureDemod (the rtl_fm shape,
 * rustradio-ui/examples/rtlsdr-fm/src/worker.rs:85-90): produces out_n-1
 * floats per channel, equal to running the two blocks back to back.  c32 only. */
int rrc_fir_c32_demod_run_batch(rrc_fir_t* h, const void* in_dev, size_t in_stride, size_t need,
                                float gain, float* out_dev, size_t out_stride, size_t out_n,
                                size_t nchan, void* stream);
/* Host-buffer form of the fused channelizer: nchan channels of n_in samples each, channel c at
 * in_host + c*n_in samples (u8 I/Q pairs after rrc_fir_set_input_u8iq: 2 bytes per sample over PCIe);
 * channel c's floor((n_in-ntaps+1)/deci) - 1 demodulated floats go to out_host + c*out_stride. */
int rrc_fir_c32_demod_run_host_batch(rrc_fir_t* h, const void* in_host, size_t n_in, size_t nchan, float gain,
                                     float* out_host, size_t out_stride, size_t* n_out_per_chan);
/* Host-buffer form: H2D -> kernel -> D2H, chunked and double-buffered, for a
 * whole stream of n_in samples; writes floor((n_in-ntaps+1)/deci) outputs
 * (0 if n_in < ntaps+deci-1) and returns the count in *n_out. */
int rrc_fir_run_host(rrc_fir_t* h, const void* in_host, size_t n_in, void* out_host, size_t* n_out);

/* ---------------------------------------------------------- FftFilter --- */
/*
 * Replaces RustFftEngine::new / Engine::run / sum_vec and the overlap
 * handling of FftFilter::work (src/fft_filter.rs:144-176, 281-287, 331-348).
 * The device computes the same linear convolution y[n] = sum_k h[k] x[n-k]
 * (zero initial state, src/fft_filter.rs:270) by overlap-SAVE with its own
 * FFT size; state carried between calls is the last ntaps-1 inputs.
 * run() accepts any n and produces exactly n outputs; the reference's count
 * rule (whole blocks of nsamples = 2*nextpow2(ntaps) - ntaps, trailing
 * partial block never flushed, :306-327) is applied by the caller / by
 * rrc_fftfilt_plan.
 * Kernel selection of run(): Complex c32 streams take fftfilt_tmh_kernel (TMA-staged input; the filter spectrum and the
 * phase-B twiddles are thread-private tables in tensor memory: RRC_FFTFILT_VARIANT=42, the default; 40 / 41 = other table
 * choices, 36 = fftfilt_tma_kernel with the spectrum half in shared memory, half from L2, 32 = the LDG kernel); u8 I/Q input
 * and real (f32) streams take fftfilt_kernel.  Every variant computes the same values (same tables, same arithmetic).
 */
typedef struct rrc_fftfilt rrc_fftfilt_t;

int rrc_fftfilt_c32_create(int device, const float* taps_c32, size_t ntaps, rrc_fftfilt_t** out);
/* FftFilterFloat (src/fft_filter.rs:365-491): real taps on a real stream.  The handle's run /
 * run_host / set_history then take f32 arrays (n, history length and counts in samples as before).
 * The reference widens to Complex, runs the complex filter and keeps .re; here two consecutive real
 * blocks ride through one complex transform as its real and imaginary parts (exact for real taps),
 * i.e. half the transforms and no widened intermediate. */
int rrc_fftfilt_f32_create(int device, const float* taps_f32, size_t ntaps, rrc_fftfilt_t** out);
int rrc_fftfilt_destroy(rrc_fftfilt_t* h);
int rrc_fftfilt_reset(rrc_fftfilt_t* h, void* stream);          /* zero the carried history */
/* Load the carried history (the ntaps-1 samples that precede the next run's input) from a
 * device buffer: the left halo of a time-segment shard, e.g. received from the neighbouring
 * GPU over NVLink.  n_samples must be ntaps-1. */
int rrc_fftfilt_set_history(rrc_fftfilt_t* h, const float* hist_dev_c32, size_t n_samples, void* stream);
/* Zero-copy form: the NEXT run / decim_run reads its ntaps-1 sample left halo through `hist_dev_c32` itself
 * (one-shot; e.g. the IPC- or peer-mapped tail of the left neighbour's input buffer: only the kernel's first
 * block touches it, over NVLink).  The pointer must stay valid until that run has completed. */
int rrc_fftfilt_set_history_ptr(rrc_fftfilt_t* h, const float* hist_dev_c32, size_t n_samples);
/* Store epilogue of rrc_fftfilt_run / decim_run / *_run_host (RRC_EPI_*, see rrc_fir_set_epilogue): Complex filters;
 * filters split into tap partitions (ntaps > 12289) take it only through the fused decimate-by-8 kernel. */
int rrc_fftfilt_set_epilogue(rrc_fftfilt_t* h, int kind, float re, float im);
/* calc_fft_size and nsamples exactly as the reference (src/fft_filter.rs:36-42,262-263). */
int rrc_fftfilt_ref_fft_size(size_t ntaps, size_t* fft_size, size_t* nsamples);
/* Device-side geometry actually used (FFT size, valid outputs per block). */
int rrc_fftfilt_geometry(const rrc_fftfilt_t* h, size_t* fft_size, size_t* valid_per_block);
/* Integer part of FftFilter::work's loop (src/fft_filter.rs:293-327) for one
 * call: with `buffered` samples already accumulated (< nsamples), an input
 * window of in_len and out_free output space, how many whole reference
 * blocks can run now (*blocks), how many input samples are taken (*consume,
 * including a final partial accumulation) and the WaitForStream that ends
 * the loop. */
int rrc_fftfilt_plan(size_t ntaps, size_t buffered, size_t in_len, size_t out_free,
                     size_t* blocks, size_t* consume, size_t* buffered_after,
                     size_t* wait_need, int* wait_on_output);
int rrc_fftfilt_run(rrc_fftfilt_t* h, const float* in_dev, size_t n, float* out_dev, void* stream);
/* Fused FftFilter -> RationalResampler(1, deci) (BASELINE config 5): writes
 * y[k*deci - phase] style decimated output without materialising y.
 * `skip` = number of filter outputs to drop before the first kept one
 * (carries the resampler counter between calls); produces
 * ceil((n - skip)/deci) outputs for n > skip.
 * Kernel selection: 2 <= deci <= 16 on a Complex filter of at most 8193 * deci taps runs as a POLYPHASE filter
 * (fftfilt_poly_kernel: deci forward transforms on the deci-times slower branch streams, one inverse per block, the sum
 * over the branches in tensor memory; clusters of 4 / 2 / 1 CTAs by what divides deci) — RRC_FFTFILT_NO_POLY=1 disables
 * it; then deci == 8 takes the folded-spectrum kernel (RRC_FFTFILT_NO_FOLD=1 disables) and everything else the plain
 * kernel with a store predicate.  Same outputs within the FftFilter tolerance on every path; one launch per call. */
int rrc_fftfilt_decim_run(rrc_fftfilt_t* h, const float* in_dev, size_t n, size_t deci, size_t skip,
                          float* out_dev, size_t* n_out, void* stream);
/* Host-buffer form for a whole stream: n_in samples in, floor(n_in/nsamples)*nsamples out. */
int rrc_fftfilt_run_host(rrc_fftfilt_t* h, const float* in_host, size_t n_in, float* out_host, size_t* n_out);

/* ------------------------------------------------------- RtlSdrDecode --- */
/*
 * SURVEY 8f rank 1 — the block that feeds the hot path in the rtl_fm chain.
 * Replaces RtlSdrDecode::work (src/rtlsdr_decode.rs:18-48): byte pairs (I, Q)
 * -> Complex((I - 127.0) * 0.008, (Q - 127.0) * 0.008); bit-exact; tags dropped.
 */
/* Integer part of work() run to its WaitForStream: with in_len_bytes readable
 * bytes and out_free output samples, consume (in_len & !1, capped by 2*out_free)
 * bytes and produce half as many samples (:23-33); then WaitForStream(src, 2)
 * or WaitForStream(dst, 1). */
int rrc_rtlsdr_decode_plan(size_t in_len_bytes, size_t out_free, size_t* consume_bytes, size_t* produce,
                           size_t* wait_need, int* wait_on_output);
/* n_bytes/2 samples from device bytes to device c32 (any alignment). */
int rrc_rtlsdr_decode_run(int device, const unsigned char* in_dev, size_t n_bytes, float* out_dev_c32, void* stream);
int rrc_rtlsdr_decode_run_host(int device, const unsigned char* in_host, size_t n_bytes, float* out_host_c32, size_t* n_out);
/* RtlSdrEncode::work (src/rtlsdr_encode.rs:22-52), the inverse wire format: Complex -> (u8 I, u8 Q) with
 * ((s / 0.008) + 127).round().clamp(0, 255) as u8 — bit-exact (f32 division and addition, round half away from
 * zero, NaN -> 0 like the saturating cast).  _plan: the reference loop run to its WaitForStream for `in_len`
 * readable samples and `out_free_bytes` writable bytes: consume samples, produce 2 bytes each; then
 * wait_on_output = 0: WaitForStream(src, 1), = 1: WaitForStream(dst, 2). */
int rrc_rtlsdr_encode_plan(size_t in_len, size_t out_free_bytes, size_t* consume, size_t* produce_bytes,
                           size_t* wait_need, int* wait_on_output);
int rrc_rtlsdr_encode_run(int device, const float* in_dev_c32, size_t n, unsigned char* out_dev, void* stream);
int rrc_rtlsdr_encode_run_host(int device, const float* in_host_c32, size_t n, unsigned char* out_host, size_t* n_out_bytes);
/* Fused form: the FIR / FftFilter kernels decode u8 I/Q pairs in their first
 * load, so RtlSdrDecode -> FirFilter<Complex> / FftFilter chains never
 * materialise the c32 stream (and *_run_host moves 2 B/sample over PCIe).
 * After set_input_u8iq(h, 1) every `in` pointer of that handle's run /
 * run_batch / decim_run / run_host calls is a 2-byte aligned array of u8 pairs;
 * counts and strides stay in samples; outputs and carried history stay c32.
 * Results equal RtlSdrDecode followed by the c32 filter bit for bit. */
int rrc_fir_set_input_u8iq(rrc_fir_t* h, int on);
int rrc_fftfilt_set_input_u8iq(rrc_fftfilt_t* h, int on);

/* Host-buffer form of FftFilter -> RationalResampler(1, deci) for a whole stream (config 5 end to
 * end): floor(n_in/nsamples)*nsamples filter outputs, every deci-th kept starting with the first. */
int rrc_fftfilt_decim_run_host(rrc_fftfilt_t* h, const float* in_host, size_t n_in, size_t deci, float* out_host, size_t* n_out);

/* ------------------------------------------------- FftStream / Fft --- */
/*
 * SURVEY 8f rank 2.  Replaces the compute of FftStream::work (src/fft_stream.rs:71-117: every
 * `size` samples become their unnormalised forward DFT, rustfft process()) and Fft::process_one
 * (src/fft.rs:31-35).  Device sizes: powers of two up to 16384 (others: RRC_ERR_UNSUPPORTED;
 * size 0: RRC_ERR_INVALID like the reference's assert / Err, :42 / src/fft.rs:25-27).
 */
typedef struct rrc_fft rrc_fft_t;
int rrc_fft_c32_create(int device, size_t size, rrc_fft_t** out);
int rrc_fft_destroy(rrc_fft_t* h);
int rrc_fft_size(const rrc_fft_t* h, size_t* size);
/* nframes consecutive frames of `size` c32 samples; in place allowed. */
int rrc_fft_run(rrc_fft_t* h, const float* in_dev_c32, size_t nframes, float* out_dev_c32, void* stream);
/* Integer part of FftStream::work (:73-84): *len samples (a multiple of size) are transformed
 * now, or len = 0 and WaitForStream(src|dst, size). */
int rrc_fftstream_plan(size_t size, size_t in_len, size_t out_free, size_t* len, size_t* wait_need, int* wait_on_output);
int rrc_fft_run_host(rrc_fft_t* h, const float* in_host, size_t n_in, float* out_host, size_t* n_out);

/* -------------------------------------------------- RationalResampler --- */
/*
 * Replaces RationalResampler::new / work (src/rational_resampler.rs:125-206):
 * counter-driven sample-and-hold/drop, out[k] = in[floor((k*deci - c0)/interp)].
 * No filtering, no floating point; bit-exact for any element size.
 */
typedef struct rrc_resampler rrc_resampler_t;

/* elem_size in {1,2,4,8,16}.  interp == 0 or deci == 0 -> RRC_ERR_INVALID
 * (the reference returns Err, :130-135).  gcd-reduced like :136-138. */
int rrc_resampler_create(int device, size_t elem_size, size_t interp, size_t deci, rrc_resampler_t** out);
int rrc_resampler_destroy(rrc_resampler_t* h);
int rrc_resampler_reset(rrc_resampler_t* h);
/* Set the carried state (src/rational_resampler.rs:101-105).  A time-segment shard that owns outputs
 * [k_lo, k_hi) starts at input s = floor(k_lo*deci/interp) with counter = s*interp - k_lo*deci (<= 0) and no
 * pending sample.  pending_host != NULL (elem_size bytes, host memory) makes that sample the pending one;
 * counter must then be > 0.  interp/deci are the gcd-reduced values (rrc_resampler_state). */
int rrc_resampler_set_state(rrc_resampler_t* h, int64_t counter, const void* pending_host);
/* State inspection (counter <= 0 between calls unless a sample is pending). */
int rrc_resampler_state(const rrc_resampler_t* h, int64_t* interp, int64_t* deci, int64_t* counter, int* has_pending);
/* One work() call on an input window of n_in and an output window of out_cap
 * samples.  *wait_on_output: 1 = WaitForStream(dst,1), 0 = WaitForStream(src,1). */
int rrc_resampler_run(rrc_resampler_t* h, const void* in_dev, size_t n_in, void* out_dev, size_t out_cap,
                      size_t* consumed, size_t* produced, int* wait_on_output, void* stream);
int rrc_resampler_run_host(rrc_resampler_t* h, const void* in_host, size_t n_in, void* out_host, size_t out_cap,
                           size_t* consumed, size_t* produced);

/* ---------------------------------------------------- QuadratureDemod --- */
/*
 * Replaces the compute of QuadratureDemod::work (src/quadrature_demod.rs:56-111,
 * libm branch): out[t] = gain * atan2(Im, Re) of conj(in[t]) * in[t+1],
 * t < n_in - 1.  The caller keeps the 1-sample history by consuming n_in - 1.
 */
int rrc_quad_demod_run(int device, const float* in_dev_c32, size_t n_in, float gain, float* out_dev, void* stream);
int rrc_quad_demod_run_batch(int device, const float* in_dev_c32, size_t in_stride, size_t n_in, float gain,
                             float* out_dev, size_t out_stride, size_t nchan, void* stream);
int rrc_quad_demod_run_host(int device, const float* in_host_c32, size_t n_in, float gain, float* out_host);

/* ------------------------------------------------------------ Hilbert --- */
/*
 * SURVEY 8f rank 3a.  Replaces WindowType::make_window (src/window.rs:63-185), fir::hilbert
 * (src/fir.rs:660-680) and the compute of Hilbert::work (src/hilbert.rs:72-128): with
 * z = [0]*ntaps ++ x (the block's carried history starts as zeros, :52),
 *   out[i] = Complex(z[i + ntaps/2], sum_j z[i+j] * taps[ntaps-1-j]),  i < n   (:106-114)
 * n samples in -> n samples out, chunking independent; the handle carries the ntaps-sample history.
 */
#define RRC_WINDOW_HAMMING          0   /* a0 = 25/46 (src/window.rs:36-37) */
#define RRC_WINDOW_BLACKMAN         1
#define RRC_WINDOW_BLACKMAN_HARRIS  2
#define RRC_WINDOW_HAMMING_PARM     3   /* HammingParm(parm) */
int rrc_make_window(int window_type, float parm, size_t ntaps, float* window_out);           /* host */
int rrc_hilbert_taps(const float* window, size_t ntaps, float* taps_out);                    /* host */
typedef struct rrc_hilbert rrc_hilbert_t;
/* taps in caller order (fir::hilbert output); ntaps must be odd and > 1 (RRC_ERR_INVALID, the
 * reference asserts, :44-47); at most 8191 taps (RRC_ERR_UNSUPPORTED beyond). */
int rrc_hilbert_create(int device, const float* taps, size_t ntaps, rrc_hilbert_t** out);
int rrc_hilbert_destroy(rrc_hilbert_t* h);
int rrc_hilbert_reset(rrc_hilbert_t* h, void* stream);                                       /* history <- zeros */
int rrc_hilbert_run(rrc_hilbert_t* h, const float* in_dev, size_t n, float* out_dev_c32, void* stream);

/* ------------------------------------------- sample-wise neighbours --- */
/*
 * the `sync` blocks that sit between the filters in real chains.  Each `n` is in
 * samples; in place is allowed for the maps.  MultiplyConst / AddConst / ComplexToMag2 / Tee are
 * bit-exact (separately rounded f32 operations, num-complex multiplication order).
 *   MultiplyConst<T>::process_sync  x * val        src/multiply_const.rs:16-23
 *   AddConst<T>::process_sync       x + val        src/add_const.rs:36-44
 *   ComplexToMag2::process_sync     norm_sqr()     src/complex_to_mag2.rs:17-20
 *   Tee<T>::process_sync            (s, s)         src/tee.rs:20-24
 */
int rrc_multiply_const_f32_run(int device, const float* in_dev, size_t n, float val, float* out_dev, void* stream);
int rrc_multiply_const_c32_run(int device, const float* in_dev_c32, size_t n, float val_re, float val_im, float* out_dev_c32, void* stream);
int rrc_add_const_f32_run(int device, const float* in_dev, size_t n, float val, float* out_dev, void* stream);
int rrc_add_const_c32_run(int device, const float* in_dev_c32, size_t n, float val_re, float val_im, float* out_dev_c32, void* stream);
int rrc_complex_to_mag2_run(int device, const float* in_dev_c32, size_t n, float* out_dev, void* stream);
int rrc_tee_run(int device, const void* in_dev, size_t nbytes, void* out1_dev, void* out2_dev, void* stream);
/*
 * IqBalance (src/iq_balance.rs:12-81): mean[n] = mean[n-1]*(1-alpha) + x[n]*alpha; out = x - mean.
 * Evaluated as a parallel affine scan (same recurrence, different association than the reference's
 * sequential f32 loop: tolerance parity, not bit-exact).  The handle carries `mean` across calls.
 */
typedef struct rrc_iq_balance rrc_iq_balance_t;
int rrc_iq_balance_alpha_from_tau(unsigned sample_rate, double tau_seconds, float* alpha);   /* with_tau, :41-57 (host) */
int rrc_iq_balance_create(int device, float alpha, rrc_iq_balance_t** out);                  /* with_alpha, :62-73; alpha clamped to [0,1] */
int rrc_iq_balance_destroy(rrc_iq_balance_t* h);
int rrc_iq_balance_reset(rrc_iq_balance_t* h, void* stream);
int rrc_iq_balance_mean(rrc_iq_balance_t* h, float* mean_re_im, void* stream);               /* D2H + sync */
int rrc_iq_balance_run(rrc_iq_balance_t* h, const float* in_dev_c32, size_t n, float* out_dev_c32, void* stream);

/* ------------------------------------------- block-level contract (rrb_) --- */
/*
 * rustradio's Block / ReadStream / WriteStream / Tag contract (src/block.rs:12-126,
 * src/stream.rs:48-339, src/nowasm/circular_buffer.rs:340-616) restated over the
 * kernels above, so a host language without the `rustradio-cuda` Rust crate can
 * still drive the blocks exactly as the reference's tests drive its own:
 * constructors take ownership of the input ReadStream and hand back the block
 * plus the output ReadStream; work() reports Again / WaitForStream(stream, need)
 * / EOF; tags travel with the samples.  Streams are double-mapped rings in
 * pageable host memory (residency 0), page-locked host memory (residency 2) or
 * device memory (residency 1, CUDA VMM) with a configurable size (the
 * reference's is fixed at 4,096,000 bytes).
 */
typedef struct rrb_rstream rrb_rstream_t;     /* ReadStream<T> */
typedef struct rrb_wstream rrb_wstream_t;     /* WriteStream<T> */
typedef struct rrb_block rrb_block_t;         /* Box<dyn Block> */

#define RRB_TAG_STRING 0
#define RRB_TAG_FLOAT  1
#define RRB_TAG_BOOL   2
#define RRB_TAG_U64    3
#define RRB_TAG_I64    4
typedef struct {
    uint64_t pos;          /* relative to the window, like Tag::pos() */
    const char* key;
    int kind;              /* RRB_TAG_* (TagValue variant, src/stream.rs:17-34) */
    const char* s;
    float f;
    int b;
    uint64_t u;
    int64_t i;
} rrb_tag_t;

#define RRB_RET_AGAIN   0
#define RRB_RET_PENDING 1
#define RRB_RET_WAIT    2   /* WaitForStream(stream_id, need) */
#define RRB_RET_EOF     3

#define RRB_HOST   0          /* pageable host ring (memfd double mapping, the reference's layout) */
#define RRB_DEVICE 1          /* device ring (CUDA VMM double mapping): chained GPU blocks never touch host memory */
#define RRB_HOST_PINNED 2     /* host ring, page-locked with cudaHostRegister: the CPU<->GPU edges of a graph are real async DMA */
#define RRB_DEFAULT_STREAM_SIZE 4096000

int rrb_stream_new(size_t elem_size, size_t bytes, int residency, int device, rrb_wstream_t** w, rrb_rstream_t** r);
/* write_buf() + fill_from_slice + produce(n, tags): writes min(n, free) samples. */
int rrb_wstream_write(rrb_wstream_t* w, const void* host_data, size_t n, const rrb_tag_t* tags, size_t ntags, size_t* written);
int rrb_wstream_free(rrb_wstream_t* w, size_t* free_samples);
int rrb_wstream_id(rrb_wstream_t* w, size_t* id);
int rrb_wstream_drop(rrb_wstream_t* w);                         /* dropping the writer is how EOF propagates */
/* read_buf(): copies up to max samples of the current window to host_out (no consume) and
 * snapshots the window's tags for rrb_rstream_tag(). */
int rrb_rstream_read(rrb_rstream_t* r, void* host_out, size_t max, size_t* window_len, size_t* ntags);
int rrb_rstream_tag(rrb_rstream_t* r, size_t index, rrb_tag_t* tag);   /* pointers valid until the next read */
int rrb_rstream_consume(rrb_rstream_t* r, size_t n);
int rrb_rstream_id(rrb_rstream_t* r, size_t* id);
int rrb_rstream_capacity(rrb_rstream_t* r, size_t* samples);
int rrb_rstream_eof(rrb_rstream_t* r, int* eof);
int rrb_rstream_drop(rrb_rstream_t* r);

/* Constructors: `src` is consumed iff the call returns RRC_OK (on any error the caller still owns it and
 * must drop it).  out_bytes/out_residency/device configure the output stream.  A device-resident `src`
 * ring must live on `device` (RRC_ERR_INVALID otherwise: cross-device chains need a host edge). */
int rrb_vector_source_new(const void* data, size_t n, size_t elem_size, uint64_t repeat,
                          size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
/* Capture ingest: FileSource<T>::builder(path).repeat(r).build() (src/file_source.rs:11-153) —
 * raw little-endian samples of elem_size bytes (8 = a cf32 capture); whole samples only; work() returns Again /
 * Pending / WaitForStream(dst,1) / EOF exactly like the reference.  repeat = number of passes (Repeat::finite),
 * UINT64_MAX = Repeat::infinite.  With a DEVICE output ring the bytes go file -> pinned staging -> HBM. */
int rrb_file_source_new(const char* path, size_t elem_size, uint64_t repeat,
                        size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
/* SigMFSource<T>::builder(path).sample_rate().repeat().ignore_type_error().build() (src/sigmf.rs:229-613): a SigMF
 * Archive (tar with X.sigmf-meta / X.sigmf-data) when `path` exists, else the Recording files path-meta / path-data.
 * type_string = the reference's Type::type_string() ("cf32", "rf32", "ru8", "ri32", "ci32"); core:datatype must be
 * type_string + "_le" unless ignore_type_error.  samp_rate < 0 = None; a rate in the metadata must equal it. */
int rrb_sigmf_source_new(const char* path, size_t elem_size, const char* type_string, double samp_rate, int ignore_type_error,
                         uint64_t repeat, size_t out_bytes, int out_residency, int device,
                         rrb_block_t** blk, rrb_rstream_t** out, double* sample_rate_out, int* has_sample_rate);
int rrb_fir_filter_new(rrb_rstream_t* src, int cplx, const float* taps, size_t ntaps, size_t deci,
                       int translate, float samp_rate, float freq, unsigned flags,
                       size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
int rrb_fft_filter_new(rrb_rstream_t* src, const float* taps_c32, size_t ntaps,
                       size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
int rrb_fft_filter_float_new(rrb_rstream_t* src, const float* taps, size_t ntaps,
                             size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
int rrb_rational_resampler_new(rrb_rstream_t* src, size_t interp, size_t deci,
                               size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
int rrb_quadrature_demod_new(rrb_rstream_t* src, float gain,
                             size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
/* FftStream::new(src, size) -> (Self, ReadStream<Complex>)  (src/fft_stream.rs:38-60); frame tags
 * "FftStream::size" / "FftStream::frame" as :95-106. */
int rrb_fft_stream_new(rrb_rstream_t* src, size_t size,
                       size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
/* RtlSdrDecode::new(src: ReadStream<u8>) -> (Self, ReadStream<Complex>)  (src/rtlsdr_decode.rs:9-16) */
int rrb_rtlsdr_decode_new(rrb_rstream_t* src,
                          size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
/* RtlSdrEncode::new(src: ReadStream<Complex>) -> (Self, ReadStream<u8>)  (src/rtlsdr_encode.rs:12-20) */
int rrb_rtlsdr_encode_new(rrb_rstream_t* src,
                          size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
/* Capture egress: FileSink<T>::builder(path).mode(m).flush(f).build(src) (src/file_sink.rs:11-160) — raw little-endian
 * samples of the stream's element size.  mode: RRB_FILE_CREATE fails when the file exists, RRB_FILE_OVERWRITE truncates,
 * RRB_FILE_APPEND appends.  work(): everything readable is written, Again; WaitForStream(src, 1) on an empty stream.
 * A DEVICE input ring is read through a pinned staging buffer.  A sink has no output stream. */
#define RRB_FILE_CREATE    0
#define RRB_FILE_OVERWRITE 1
#define RRB_FILE_APPEND    2
int rrb_file_sink_new(rrb_rstream_t* src, const char* path, int mode, int flush, int device, rrb_block_t** blk);
/* Hilbert::new(src: ReadStream<Float>, ntaps, &WindowType) -> (Self, ReadStream<Complex>)  (src/hilbert.rs:35-60) */
int rrb_hilbert_new(rrb_rstream_t* src, size_t ntaps, int window_type, float window_parm,
                    size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
/* `sync` blocks (rustradio_macros_code/src/lib.rs:436-514): loop { n = min(in, out space); map; consume n;
 * produce n with the input tags at unchanged positions } until WaitForStream(src|dst, 1).
 * cplx selects T = Complex (val = re + i im) or T = Float (val = re). */
int rrb_multiply_const_new(rrb_rstream_t* src, int cplx, float val_re, float val_im,
                           size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
int rrb_add_const_new(rrb_rstream_t* src, int cplx, float val_re, float val_im,
                      size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
int rrb_complex_to_mag2_new(rrb_rstream_t* src,
                            size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
/* Tee::new(src) -> (Self, ReadStream<T>, ReadStream<T>)  (src/tee.rs:9-18); tags go to both outputs. */
int rrb_tee_new(rrb_rstream_t* src,
                size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out1, rrb_rstream_t** out2);
/* IqBalance::with_alpha(src, alpha) (src/iq_balance.rs:62-73); use rrc_iq_balance_alpha_from_tau for new()/with_tau(). */
int rrb_iq_balance_new(rrb_rstream_t* src, float alpha,
                       size_t out_bytes, int out_residency, int device, rrb_block_t** blk, rrb_rstream_t** out);
int rrb_block_work(rrb_block_t* b, int* kind, size_t* stream_id, size_t* need);
int rrb_block_eof(rrb_block_t* b, int* eof);
const char* rrb_block_name(rrb_block_t* b);
int rrb_block_drop(rrb_block_t* b);
/* Graph::run (src/graph.rs:99-173) over the given blocks, single threaded, round robin. */
int rrb_graph_run(rrb_block_t** blocks, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* RUSTRADIO_CUDA_H */
