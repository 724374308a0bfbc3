//! rustradio-cuda: GPU drop-ins for `rustradio::blocks::{FirFilter, FftFilter, RationalResampler,
//! QuadratureDemod}` with the same constructor signatures and the same `Block::work()` contract.
//! UNCOMPILED in this repository's environment (no Rust toolchain) — see INTEGRATION.md.
pub mod ffi;
pub mod blocks;
pub use blocks::{
    CudaFftFilter, CudaFftFilterFloat, CudaFirFilter, CudaFirFilterBuilder, CudaQuadratureDemod, CudaRationalResampler,
    CudaRationalResamplerBuilder, CudaRtlSdrDecode, GpuSample,
};
