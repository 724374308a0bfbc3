"""rustradio_b200 — B200-native filtering hot path for rustradio.

The product is the C-ABI shared library `librustradio_cuda.so`
(include/rustradio_cuda.h; sources in rustradio_b200/csrc).  This Python
package is only a ctypes binding of that ABI for the test-suite and bench.py;
it contains no compute and no CPU fallback: every call goes to the CUDA
library and raises `RrcError` if the library or a CUDA device is missing.
"""
from .api import (  # noqa: F401
    RrcError, lib, library_path, build_library, device_count,
    DeviceBuffer, PinnedBuffer, Fir, FftFilt, Resampler, quad_demod, quad_demod_host,
    rtlsdr_decode, rtlsdr_decode_host, rtlsdr_decode_plan, rtlsdr_encode, rtlsdr_encode_host, rtlsdr_encode_plan, Fft, fftstream_plan,
    synth_f32, launch_count, fir_plan, fftfilt_plan, fftfilt_ref_fft_size,
    Event, stream_sync, device_sync, event_wait, device_numa_node, peer_enable, ipc_export, IpcMapping,
    Hilbert, make_window, hilbert_taps, multiply_const, add_const, complex_to_mag2, tee, IqBalance,
    iq_balance_alpha_from_tau, WINDOW_HAMMING, WINDOW_BLACKMAN, WINDOW_BLACKMAN_HARRIS, WINDOW_HAMMING_PARM,
    RRC_FIR_NO_REAL_TAP_FASTPATH, RRC_FIR_FORCE_GENERIC, RRC_FIR_NO_TENSOR,
    EPI_NONE, EPI_MULTIPLY_CONST, EPI_ADD_CONST, EPI_MAG2,
)
