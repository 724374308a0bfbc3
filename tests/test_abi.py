"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every
symbol include/rustradio_cuda.h declares, fails loudly (no CPU fallback) when a
compute entry point is called without a CUDA device, and its host-side integer
logic (the count rules of FirFilter::work / FftFilter::work) equals the
restated reference."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import rustradio_b200 as R
from oracle import blockmodel as B
from oracle import oracle as O

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "rustradio_cuda.h").read_text()


def declared_symbols():
    return sorted(set(re.findall(r"\b(rr[cb]_[a-z0-9_]+)\s*\(", HEADER)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(str(R.library_path()))
    names = declared_symbols()
    assert len(names) > 50
    for n in names:
        assert hasattr(lib, n), f"{n} declared in rustradio_cuda.h but not exported"
    # and the Python binding covers the same set
    from rustradio_b200 import api, blocks
    assert set(api.exported_symbols()) | set(blocks.exported_symbols()) == set(names)


def test_header_cites_reference_for_each_group():
    for cite in ("src/fir.rs", "src/fft_filter.rs", "src/rational_resampler.rs", "src/quadrature_demod.rs"):
        assert cite in HEADER


def _no_gpu():
    try:
        return R.device_count() == 0
    except R.RrcError:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="box has a GPU")
def test_no_cpu_fallback_without_a_device():
    taps = np.ones(4, np.complex64)
    with pytest.raises(R.RrcError):
        R.Fir(taps)
    with pytest.raises(R.RrcError):
        R.FftFilt(taps)
    with pytest.raises(R.RrcError):
        R.Resampler(4, 1, 2)
    with pytest.raises(R.RrcError):
        R.quad_demod_host(np.ones(8, np.complex64))


def test_product_package_does_not_import_the_oracle():
    for p in (ROOT / "rustradio_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".hpp", ".cpp", ".h", ".rs"):
            txt = p.read_text(errors="replace")
            assert "oracle" not in txt.lower() or p.name in ("common.cuh", "runtime.cu") and "rr_oracle.c" in txt, p


def test_invalid_arguments_are_errors():
    assert R.lib().rrc_fir_plan(0, 1, 10, 10, *[C.byref(C.c_size_t()) for _ in range(4)], C.byref(C.c_int())) == -1
    h = C.c_void_p()
    assert R.lib().rrc_resampler_create(0, 4, 0, 1, C.byref(h)) == -1
    assert b"interp 0" in R.lib().rrc_last_error()
    assert R.lib().rrc_resampler_create(0, 4, 1, 0, C.byref(h)) == -1
    assert R.lib().rrc_resampler_create(0, 3, 1, 1, C.byref(h)) == -1
    assert R.lib().rrc_fir_c32_create(0, None, 0, 1, 0, C.byref(h)) == -1
    t = np.ones(2, np.float32)
    assert R.lib().rrc_fir_c32_create(0, t.ctypes.data, 1, 0, 0, C.byref(h)) == -1
    assert R.lib().rrc_fftfilt_c32_create(0, None, 0, C.byref(h)) == -1


@pytest.mark.parametrize("ntaps,deci", [(1, 1), (1, 5), (2, 3), (64, 1), (255, 10), (3, 18)])
def test_fir_plan_equals_restated_work(ntaps, deci):
    """rrc_fir_plan vs the block model's FirFilter::work (src/fir.rs:496-525) on random windows."""
    rng = np.random.default_rng(ntaps * 100 + deci)
    for _ in range(200):
        in_len = int(rng.integers(0, 4 * (ntaps + deci)))
        out_free = int(rng.integers(0, 6))
        consume, need, out_n, wait_need, wait_out = R.fir_plan(ntaps, deci, in_len, out_free)
        src = B.Stream(np.complex64, 8 * 4096)
        w = src.write_buf()
        w[:in_len] = 1
        src.produce(in_len)
        blk = B.FirFilter(src, np.ones(ntaps, np.complex64), deci, stream_bytes=8 * 4096)
        # occupy the output so that exactly out_free samples are free
        blk.out.produce(blk.out.cap - out_free)
        ret = blk.work()
        if ret.kind == B.AGAIN:
            assert consume == in_len - src.used and consume > 0
            assert out_n == out_free - blk.out.free()
            assert need == consume + ntaps - 1 and need <= in_len
        else:
            assert consume == 0 and wait_need == ret.need
            assert wait_out == (1 if ret.stream is blk.out else 0)


@pytest.mark.parametrize("ntaps", [1, 3, 193, 300])
def test_fftfilt_plan_equals_restated_work(ntaps):
    """rrc_fftfilt_plan vs the block model's FftFilter::work loop (src/fft_filter.rs:293-352)."""
    rng = np.random.default_rng(ntaps)
    fft_size, S = R.fftfilt_ref_fft_size(ntaps)
    assert fft_size == O.calc_fft_size(ntaps) and S == fft_size - ntaps
    cap_bytes = 8 * 8192
    for _ in range(60):
        src = B.Stream(np.complex64, cap_bytes)
        blk = B.FftFilter(src, np.ones(ntaps, np.complex64), stream_bytes=cap_bytes)
        buffered = 0
        for _ in range(4):
            in_len = int(rng.integers(0, min(src.free(), 3 * S + 5) + 1))
            src.write_buf()[:in_len] = 0
            src.produce(in_len)
            out_free = int(rng.integers(0, blk.out.free() + 1))
            pad = blk.out.free() - out_free
            blk.out.produce(pad)
            in_window = src.used
            blocks, consume, buffered_after, wait_need, wait_out = R.fftfilt_plan(ntaps, buffered, in_window, out_free)
            before_out = blk.out.used
            ret = blk.work()
            assert ret.kind == B.WAIT
            assert consume == in_window - src.used
            assert blocks * S == blk.out.used - before_out
            assert buffered_after == len(blk.buf)
            assert wait_need == ret.need and wait_out == (1 if ret.stream is blk.out else 0)
            buffered = buffered_after
            blk.out.consume(blk.out.used)


def test_rust_ffi_declares_header_symbols():
    """The (uncompiled) Rust binding is GENERATED from the header (tools/gen_rust_ffi.py): it declares every
    rrc_* / rrb_* entry point with the header's arity, and the committed file is not stale."""
    import subprocess
    import sys
    ffi = (ROOT / "rustradio_b200" / "rust" / "rustradio-cuda" / "src" / "ffi.rs").read_text()
    decl = dict(re.findall(r"pub fn (rr[cb]_[a-z0-9_]+)\s*\(([^;]*?)\)\s*->", ffi, flags=re.S))
    assert set(decl) == set(declared_symbols()), set(declared_symbols()) ^ set(decl)
    code = re.sub(r"/\*.*?\*/", " ", HEADER, flags=re.S)                 # prototypes only, no comments
    for name, args in decl.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", code, flags=re.S)
        assert m, f"{name} in ffi.rs but not in rustradio_cuda.h"
        n_rust = 0 if not args.strip() else len([a for a in args.split(",") if a.strip()])
        c_args = m.group(1).strip()
        n_c = 0 if c_args in ("", "void") else len([a for a in c_args.split(",") if a.strip()])
        assert n_rust == n_c, f"{name}: {n_rust} args in ffi.rs vs {n_c} in the header"
    assert subprocess.run([sys.executable, str(ROOT / "tools" / "gen_rust_ffi.py"), "--check"]).returncode == 0, \
        "ffi.rs is stale: run python tools/gen_rust_ffi.py"


def test_rust_blocks_cover_the_reference_constructors():
    """The Rust side exposes the reference's constructor surface for the path:
    builder + deci + translate, FirFilter<Float>, FftFilterFloat, the resampler's typestate builder and its
    pending-aware eof()."""
    rs = (ROOT / "rustradio_b200" / "rust" / "rustradio-cuda" / "src" / "blocks.rs").read_text()
    for needle in ("pub fn builder(taps: impl Into<Vec<T>>) -> CudaFirFilterBuilder<T>", "pub fn deci(mut self, deci: usize) -> Self",
                   "pub fn translate(mut self, samp_rate: Float, freq: Float) -> Self", "impl GpuSample for Float",
                   "pub type CudaFftFilterFloat = CudaFftFilterT<Float>;", "pub struct CudaRationalResamplerBuilderBoth<T>",
                   "pending == 0 && self.src.eof()", "pub struct CudaQuadratureDemod", "pub struct CudaRtlSdrDecode",
                   "pub struct CudaRtlSdrEncode"):
        assert needle in rs, needle
    # every ffi function the blocks call exists in the generated binding
    ffi = (ROOT / "rustradio_b200" / "rust" / "rustradio-cuda" / "src" / "ffi.rs").read_text()
    for fn in set(re.findall(r"ffi::(rr[cb]_[a-z0-9_]+)\(", rs)):
        assert f"pub fn {fn}(" in ffi, fn


def test_rust_build_lists_every_source():
    """build.rs compiles the same CUDA sources as csrc/Makefile."""
    mk = (ROOT / "rustradio_b200" / "csrc" / "Makefile").read_text()
    srcs = re.search(r"^SRCS := (.*)$", mk, flags=re.M).group(1).split()
    br = (ROOT / "rustradio_b200" / "rust" / "rustradio-cuda" / "build.rs").read_text()
    for s_ in srcs:
        assert f'"{s_}"' in br, s_
