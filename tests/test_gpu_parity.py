"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star): FIR / FftFilter rel-RMS <= 1e-5 against the
f64 truth of the same f32 inputs; resampler bit-exact; demod <= 1e-4 rad;
sample counts exactly the reference's.  The bit-faithful f32 oracle error is
printed next to ours as a second opinion.
"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

REL_RMS_BAR = 1e-5
DEMOD_BAR = 1e-4


@pytest.fixture(scope="module")
def R():
    import rustradio_b200 as R
    assert R.device_count() >= 1
    return R


def cplx_taps(n, seed=1):
    r = np.random.default_rng(seed)
    return ((r.standard_normal(n) + 1j * r.standard_normal(n)) / np.sqrt(n)).astype(np.complex64)


# ------------------------------------------------------------------ FIR ---
def test_fir_kat_test_complex(R):
    """reference src/fir.rs:921-950"""
    x = np.array([1, 2, 3 + .2j, 4.1, 5, 6 + .2j], np.complex64)
    taps = np.array([.1, 1, .2j], np.complex64)
    assert np.allclose(R.Fir(taps).filter(x), [2.3 + .22j, 3.41 + .6j, 4.56 + .6j, 5.6 + .84j], atol=1e-6)
    assert np.allclose(R.Fir(taps, deci=2).filter(x), [2.3 + .22j, 4.56 + .6j], atol=1e-6)


@pytest.mark.parametrize("ntaps,deci,n", [
    (1, 1, 1000), (2, 1, 5000), (3, 2, 4099), (64, 1, 100_000), (247, 1, 512_000), (255, 10, 240_000),
    (33, 7, 50_001), (100, 100, 30_000), (5, 64, 100_000), (1025, 3, 70_000), (64, 1, 63), (64, 2, 64)])
@pytest.mark.parametrize("kind", ["c32_ctaps", "c32_rtaps", "f32"])
def test_fir_matches_oracle(R, ntaps, deci, n, kind):
    if kind == "f32":
        x = O.synth_f32(11, 0, n)
        taps = np.random.default_rng(ntaps).standard_normal(ntaps).astype(np.float32) / np.sqrt(ntaps)
    else:
        x = O.synth_c32(12, 0, n)
        taps = cplx_taps(ntaps, ntaps) if kind == "c32_ctaps" else O.low_pass_n(1.0, 0.1, ntaps).astype(np.complex64)
    f = R.Fir(taps, deci=deci)
    assert f.uses_real_taps == (kind == "c32_rtaps")
    y = f.filter(x)
    truth = O.fir(x, taps, deci, f64=True)
    assert len(y) == len(truth) == O.fir_out_count(n, ntaps, deci)
    if len(y):
        e = O.rel_rms(y, truth)
        e_ref = O.rel_rms(O.fir(x, taps, deci), truth)
        print(f"fir {kind} T={ntaps} D={deci}: gpu {e:.2e}  f32-oracle {e_ref:.2e}")
        assert e <= REL_RMS_BAR


@pytest.mark.parametrize("flags_name", ["RRC_FIR_FORCE_GENERIC", "RRC_FIR_NO_REAL_TAP_FASTPATH", "RRC_FIR_NO_TENSOR"])
def test_fir_alternate_paths(R, flags_name):
    x = O.synth_c32(13, 0, 20_000)
    taps = O.low_pass_n(1.0, 0.1, 101).astype(np.complex64)
    f = R.Fir(taps, deci=3, flags=getattr(R, flags_name))
    assert not (flags_name == "RRC_FIR_NO_REAL_TAP_FASTPATH" and f.uses_real_taps)
    assert not (flags_name != "RRC_FIR_FORCE_GENERIC" and f.uses_tensor_cores)
    assert O.rel_rms(f.filter(x), O.fir(x, taps, 3, f64=True)) <= REL_RMS_BAR


@pytest.mark.parametrize("ntile,nld,nm", [(1, 14, 0), (2, 14, 0), (4, 14, 0), (1, 9, 0), (2, 9, 1), (4, 9, 0)])
@pytest.mark.parametrize("ntaps,deci,n,nchan", [(64, 1, 70_001, 1), (255, 10, 61_237, 3), (31, 4, 9_000, 2), (16, 3, 40_000, 1),
                                                (200, 2, 33_333, 2)])
def test_fir_tensor_core_geometries(R, monkeypatch, ntile, nld, nm, ntaps, deci, n, nchan):
    """fir_tc_kernel: every (block-row width, register-tile size) instantiation and several tile heights, ragged
    tails, channel strides that leave odd channels 8-byte aligned only, plain and fused-demod epilogues; the FP32
    kernels as the second opinion.  RRC_FIR_TENSOR=2 takes the tensor path for shapes the planner would leave on FP32."""
    monkeypatch.setenv("RRC_FIR_TENSOR", "2")
    monkeypatch.setenv("RRC_FIR_TC1", "0")             # the generic kernel also for deci == 1
    monkeypatch.setenv("RRC_FIR_TC_NTILE", str(ntile))
    monkeypatch.setenv("RRC_FIR_TC_NLD", str(nld))
    if nm:
        monkeypatch.setenv("RRC_FIR_TC_NM", str(nm))
    taps = O.low_pass_n(1.0, 0.4 / deci, ntaps).astype(np.complex64)
    f = R.Fir(taps, deci=deci)
    f32 = R.Fir(taps, deci=deci, flags=R.RRC_FIR_NO_TENSOR)
    assert f.uses_real_taps and not f32.uses_tensor_cores
    if not f.uses_tensor_cores:
        assert (ntaps, deci) != (64, 1)
        pytest.skip("no warp tile of this geometry fits the register staging")
    stride = n + 1 if n % 2 == 0 else n            # odd stride: channel 1 starts 8 bytes off a 16-byte boundary
    xs = np.zeros((nchan, stride), np.complex64)
    for c in range(nchan):
        xs[c, :n] = O.synth_c32(200 + c, 0, n) * 0.5 + np.exp(2j * np.pi * 0.013 * (c + 1) * np.arange(n)).astype(np.complex64)
    out_n = f.out_count(n)
    need = (out_n - 1) * deci + ntaps
    din = R.DeviceBuffer.from_numpy(xs)
    ostride = out_n + 1 - (out_n % 2)              # odd output stride as well
    for filt in (f, f32):
        dy = R.DeviceBuffer(nchan * ostride * 8)
        filt.run_batch(din, stride, need, dy, ostride, out_n, nchan)
        y = dy.download(np.complex64, nchan * ostride).reshape(nchan, ostride)[:, :out_n]
        dd = R.DeviceBuffer(nchan * ostride * 4)
        filt.demod_run_batch(din, stride, need, 0.7, dd, ostride, out_n, nchan)
        d = dd.download(np.float32, nchan * ostride).reshape(nchan, ostride)[:, :out_n - 1]
        for c in range(nchan):
            truth = O.fir(xs[c, :n], taps, deci, f64=True)
            e = O.rel_rms(y[c], truth)
            print(f"fir_tc ntile={ntile} nld={nld} nm={nm} T={ntaps} D={deci} ch{c} {'tensor' if filt is f else 'fp32'}: {e:.2e}")
            assert e <= REL_RMS_BAR
            want = np.angle(truth[1:] * np.conj(truth[:-1]))
            assert O.max_angle_err(d[c] / 0.7, want) <= DEMOD_BAR


@pytest.mark.parametrize("ntaps,deci,n,nchan", [
    (16, 1, 5_000, 2), (33, 1, 70_001, 1), (64, 1, 262_144, 1), (64, 1, 1_300, 5), (100, 1, 40_000, 3), (121, 1, 20_011, 2),
    (17, 1, 600, 1), (122, 1, 9_000, 2), (137, 1, 30_000, 1), (160, 1, 5_555, 3), (169, 1, 20_000, 1), (185, 1, 7_001, 2),
    (201, 1, 12_345, 1), (217, 1, 3_000, 2), (233, 1, 40_001, 1), (249, 1, 25_000, 2),
    (18, 2, 4_000, 2), (64, 2, 100_001, 1), (127, 2, 33_333, 3), (180, 2, 9_999, 1), (242, 2, 50_000, 2),
    (40, 4, 7_000, 1), (128, 4, 123_457, 1), (129, 4, 20_000, 3), (200, 4, 5_001, 2), (228, 4, 60_000, 1),
    (255, 1, 31_000, 1), (270, 1, 8_000, 2), (290, 1, 22_222, 1), (313, 1, 10_000, 1), (255, 2, 40_000, 2), (306, 2, 9_000, 1),
    (255, 4, 70_001, 1), (292, 4, 12_000, 2), (128, 8, 99_999, 1), (255, 8, 40_000, 2), (264, 8, 7_000, 1), (60, 8, 3_000, 2)])
def test_fir_tensor_core_walk_kernel(R, monkeypatch, ntaps, deci, n, nchan):
    """fir_tc1_kernel (deci 1, 2, 4; 7*deci + ntaps <= 320; every k-step count): plain and fused-demod epilogues, ragged
    last tiles, tiles shorter than one warp tile, odd channel strides (8-byte aligned channels take the scalar loads/stores)."""
    if ntaps < 32 * deci:
        monkeypatch.setenv("RRC_FIR_TENSOR", "2")      # the planner leaves ntaps < 32*deci on the FP32 kernels
    taps = O.low_pass_n(1.0, 0.2 / deci, ntaps).astype(np.complex64)
    f = R.Fir(taps, deci=deci)
    assert f.uses_tensor_cores
    stride = n + 1 if n % 2 == 0 else n
    xs = np.zeros((nchan, stride), np.complex64)
    for c in range(nchan):
        xs[c, :n] = O.synth_c32(300 + c, 0, n) * 0.5 + np.exp(2j * np.pi * 0.013 / deci * (c + 1) * np.arange(n)).astype(np.complex64)
    out_n = f.out_count(n)
    need = (out_n - 1) * deci + ntaps
    din = R.DeviceBuffer.from_numpy(xs)
    ostride = out_n + 1 - (out_n % 2)
    dy = R.DeviceBuffer(nchan * ostride * 8)
    f.run_batch(din, stride, need, dy, ostride, out_n, nchan)
    y = dy.download(np.complex64, nchan * ostride).reshape(nchan, ostride)[:, :out_n]
    dd = R.DeviceBuffer(nchan * ostride * 4)
    f.demod_run_batch(din, stride, need, 0.7, dd, ostride, out_n, nchan)
    d = dd.download(np.float32, nchan * ostride).reshape(nchan, ostride)[:, :out_n - 1]
    for c in range(nchan):
        truth = O.fir(xs[c, :n], taps, deci, f64=True)
        e, e_ref = O.rel_rms(y[c], truth), O.rel_rms(O.fir(xs[c, :n], taps, deci), truth)
        print(f"fir_tc1 T={ntaps} D={deci} ch{c}: gpu {e:.2e}  f32-oracle {e_ref:.2e}")
        assert e <= 2e-6
        assert O.max_angle_err(d[c] / 0.7, np.angle(truth[1:] * np.conj(truth[:-1]))) <= DEMOD_BAR


@pytest.mark.parametrize("ntaps,n,nchan", [
    (64, 262_144, 1), (65, 8_256, 1), (64, 64, 1), (16, 5_000, 2), (33, 70_001, 1), (64, 1_300, 5), (48, 16_415, 3),
    (65, 3 * 8192 + 64, 2), (17, 600, 1), (40, 100_000, 2), (64, 8192 * 171 + 77, 2), (33, 8192 * 300 + 5000, 1)])
@pytest.mark.parametrize("rows", [32, 64])
def test_fir_tcgen05_kernel(R, monkeypatch, ntaps, n, nchan, rows):
    """fir_tc5_kernel (tcgen05.mma, taps and accumulators in TMEM; c32 samples, real taps, deci 1, ntaps <= 65), forced
    for every launch size with RRC_FIR_TCGEN05=2 (by default launches below 3 tiles of 8192 outputs per SM stay on fir_tc1_kernel; tiles are 4096 outputs below 12 per SM, 8192 above):
    ragged last tiles, tiles shorter than one 8192-output CTA tile, odd channel strides (8-byte aligned channels take the
    scalar loads), several k-step counts, more and fewer tiles than CTAs.  Same bar as the mma.sync kernels."""
    monkeypatch.setenv("RRC_FIR_TCGEN05", "2")
    monkeypatch.setenv("RRC_FIR_TC5_NR", str(rows))                # both tile heights at every size (default: by launch size)
    taps = O.low_pass_n(1.0, 0.2, ntaps).astype(np.complex64)
    f = R.Fir(taps)
    assert f.uses_tensor_cores and "fir_tc5_kernel" in f.kernel_name
    stride = n + 1 if n % 2 == 0 else n
    xs = np.zeros((nchan, stride), np.complex64)
    for c in range(nchan):
        xs[c, :n] = O.synth_c32(400 + c, 0, n) * 0.5 + np.exp(2j * np.pi * 0.013 * (c + 1) * np.arange(n)).astype(np.complex64)
    out_n = f.out_count(n)
    need = out_n - 1 + ntaps
    din = R.DeviceBuffer.from_numpy(xs)
    ostride = out_n + 1 - (out_n % 2)
    dy = R.DeviceBuffer(nchan * ostride * 8)
    f.run_batch(din, stride, need, dy, ostride, out_n, nchan)
    assert f.kernel_name.startswith("fir_tc5_kernel") and "fir_tc1" not in f.kernel_name    # the launch took the tcgen05 kernel
    y = dy.download(np.complex64, nchan * ostride).reshape(nchan, ostride)[:, :out_n]
    for c in range(nchan):
        truth = O.fir(xs[c, :n], taps, 1, f64=True)
        e, e_ref = O.rel_rms(y[c], truth), O.rel_rms(O.fir(xs[c, :n], taps, 1), truth)
        print(f"fir_tc5 T={ntaps} ch{c}: gpu {e:.2e}  f32-oracle {e_ref:.2e}")
        assert e <= 2e-6
    # the fused demod is not implemented on this kernel: it must still be right (falls to fir_tc1_kernel)
    dd = R.DeviceBuffer(nchan * ostride * 4)
    if out_n > 1:
        f.demod_run_batch(din, stride, need, 0.7, dd, ostride, out_n, nchan)
        assert f.kernel_name.startswith("fir_tc1_kernel")
        d = dd.download(np.float32, nchan * ostride).reshape(nchan, ostride)[:, :out_n - 1]
        truth = O.fir(xs[0, :n], taps, 1, f64=True)
        assert O.max_angle_err(d[0] / 0.7, np.angle(truth[1:] * np.conj(truth[:-1]))) <= DEMOD_BAR


@pytest.mark.parametrize("ntaps,n,nchan", [
    (64, 262_144, 1), (65, 8_256, 1), (16, 5_001, 2), (33, 70_001, 1), (64, 1_301, 5), (48, 16_415, 3), (64, 8192 * 150 + 77, 2)])
def test_fir_tcgen05_kernel_f32_streams(R, monkeypatch, ntaps, n, nchan):
    """fir_tc5_kernel on FirFilter<Float> streams (one component: three MMAs per k-step, f32 loads and stores): ragged
    tiles, odd channel strides (4-byte aligned channels take the scalar loads); 64-row tiles."""
    monkeypatch.setenv("RRC_FIR_TCGEN05", "2")
    taps = O.low_pass_n(1.0, 0.2, ntaps)
    f = R.Fir(taps)
    assert f.uses_tensor_cores and not f.cplx and "fir_tc5_kernel" in f.kernel_name
    stride = n + 1 if n % 2 == 0 else n
    xs = np.zeros((nchan, stride), np.float32)
    for c in range(nchan):
        xs[c, :n] = O.synth_f32(500 + c, 0, n) * 0.5 + np.cos(2 * np.pi * 0.013 * (c + 1) * np.arange(n)).astype(np.float32)
    out_n = f.out_count(n)
    din = R.DeviceBuffer.from_numpy(xs)
    ostride = out_n + 1 - (out_n % 2)
    dy = R.DeviceBuffer(nchan * ostride * 4)
    f.run_batch(din, stride, out_n - 1 + ntaps, dy, ostride, out_n, nchan)
    assert f.kernel_name.startswith("fir_tc5_kernel")
    y = dy.download(np.float32, nchan * ostride).reshape(nchan, ostride)[:, :out_n]
    for c in range(nchan):
        e = O.rel_rms(y[c], O.fir(xs[c, :n], taps, 1, f64=True))
        print(f"fir_tc5 f32 T={ntaps} ch{c}: {e:.2e}")
        assert e <= 2e-6


@pytest.mark.parametrize("scale", [1e-20, 1.0, 3e18])
@pytest.mark.parametrize("bad", [None, np.inf, np.nan])
@pytest.mark.parametrize("rows", [32, 64])
def test_fir_tcgen05_block_scaling_and_non_finite(R, monkeypatch, scale, bad, rows):
    """The CTA-tile power-of-two scale keeps FP32-class accuracy at any signal level; one Inf / NaN sample makes the
    outputs whose window holds it non-finite plus, at most, the rest of the two 128-output block-rows whose 192-sample
    operand rows contain it (zero-padded Toeplitz taps meet it as 0 * Inf); everything else keeps the bar."""
    monkeypatch.setenv("RRC_FIR_TCGEN05", "2")
    monkeypatch.setenv("RRC_FIR_TC5_NR", str(rows))
    n, ntaps, pos = 30_000, 64, 12_345
    taps = O.low_pass_n(1.0, 0.1, ntaps).astype(np.complex64)
    x = (O.synth_c32(61, 0, n) * np.float32(scale)).astype(np.complex64)
    f = R.Fir(taps)
    assert "fir_tc5_kernel" in f.kernel_name
    if bad is not None:
        x[pos] = bad
    y = f.filter(x)
    o = np.arange(len(y))
    xz = x.copy()
    far = np.ones(len(y), bool)
    if bad is not None:
        xz[pos] = 0
        touched = (o > pos - ntaps) & (o <= pos)
        far = (o < 128 * (pos // 128 - 1)) | (o >= 128 * (pos // 128 + 1))       # block-rows of 128 outputs read 192 samples
        assert not np.isfinite(y[touched]).any()
    assert np.isfinite(y[far]).all()
    truth = O.fir(xz, taps, 1, f64=True)
    assert O.rel_rms(y[far], truth[far]) <= 2e-6


@pytest.mark.parametrize("ntaps,deci,n,nchan", [
    (32, 1, 10_000, 2), (64, 1, 300_001, 1), (65, 1, 1_500, 3), (100, 1, 2_049, 1), (247, 1, 50_000, 2), (313, 1, 70_000, 1),
    (20, 1, 3_000, 1), (64, 2, 100_003, 1), (127, 2, 9_000, 2), (255, 2, 40_000, 1), (306, 2, 8_191, 1),
    (128, 4, 200_000, 1), (255, 4, 33_333, 2), (292, 4, 5_000, 1), (50, 4, 6_000, 1), (255, 8, 100_000, 1), (130, 8, 9_001, 2)])
def test_fir_tensor_core_f32_streams(R, monkeypatch, ntaps, deci, n, nchan):
    """fir_tcf_kernel (FirFilter<Float>, deci 1/2/4, 7*deci + ntaps <= 320): against the f64 truth and the FP32 kernel,
    ragged last tiles, odd channel strides (4-byte aligned channels take the scalar loads and stores)."""
    if ntaps < 32 * deci:
        monkeypatch.setenv("RRC_FIR_TENSOR", "2")
    taps = O.low_pass_n(1.0, 0.2 / deci, ntaps)
    f = R.Fir(taps, deci=deci)
    f32 = R.Fir(taps, deci=deci, flags=R.RRC_FIR_NO_TENSOR)
    assert f.uses_tensor_cores and not f32.uses_tensor_cores and not f.cplx
    stride = n + 1 if n % 2 == 0 else n
    xs = np.zeros((nchan, stride), np.float32)
    for c in range(nchan):
        xs[c, :n] = O.synth_f32(400 + c, 0, n) * 0.5 + np.cos(2 * np.pi * 0.013 / deci * (c + 1) * np.arange(n)).astype(np.float32)
    out_n = f.out_count(n)
    need = (out_n - 1) * deci + ntaps
    din = R.DeviceBuffer.from_numpy(xs)
    ostride = out_n + 1 - (out_n % 2)
    for filt in (f, f32):
        dy = R.DeviceBuffer(nchan * ostride * 4)
        filt.run_batch(din, stride, need, dy, ostride, out_n, nchan)
        y = dy.download(np.float32, nchan * ostride).reshape(nchan, ostride)[:, :out_n]
        for c in range(nchan):
            truth = O.fir(xs[c, :n], taps, deci, f64=True)
            e = O.rel_rms(y[c], truth)
            print(f"fir_tcf T={ntaps} D={deci} ch{c} {'tensor' if filt is f else 'fp32'}: {e:.2e}")
            assert e <= 2e-6


@pytest.mark.parametrize("ntaps,deci,n,nchan", [
    (32, 1, 10_000, 2), (64, 1, 200_001, 1), (100, 1, 1_500, 3), (249, 1, 40_000, 1), (313, 1, 9_000, 2), (20, 1, 3_000, 1),
    (64, 2, 100_003, 1), (127, 2, 9_000, 3), (306, 2, 30_000, 1), (128, 4, 150_000, 1), (255, 4, 33_333, 2), (50, 4, 6_000, 1),
    (255, 8, 60_000, 1), (100, 8, 5_000, 2)])
def test_fir_tensor_core_complex_taps(R, monkeypatch, ntaps, deci, n, nchan):
    """fir_tcc_kernel (complex taps, deci 1/2/4): plain and fused-demod epilogues against the f64 truth and the FP32
    complex-tap kernel, ragged tiles, odd channel strides."""
    if ntaps < 32 * deci:
        monkeypatch.setenv("RRC_FIR_TENSOR", "2")
    taps = (O.low_pass_n(1.0, 0.2 / deci, ntaps) * np.exp(2j * np.pi * 0.07 / deci * np.arange(ntaps))).astype(np.complex64)
    f = R.Fir(taps, deci=deci)
    f32 = R.Fir(taps, deci=deci, flags=R.RRC_FIR_NO_TENSOR)
    assert f.uses_tensor_cores and not f.uses_real_taps and not f32.uses_tensor_cores
    stride = n + 1 if n % 2 == 0 else n
    xs = np.zeros((nchan, stride), np.complex64)
    for c in range(nchan):
        xs[c, :n] = O.synth_c32(500 + c, 0, n) * 0.5 + np.exp(2j * np.pi * (0.07 + 0.011 * (c + 1)) / deci * np.arange(n)).astype(np.complex64)
    out_n = f.out_count(n)
    need = (out_n - 1) * deci + ntaps
    din = R.DeviceBuffer.from_numpy(xs)
    ostride = out_n + 1 - (out_n % 2)
    for filt in (f, f32):
        dy = R.DeviceBuffer(nchan * ostride * 8)
        filt.run_batch(din, stride, need, dy, ostride, out_n, nchan)
        y = dy.download(np.complex64, nchan * ostride).reshape(nchan, ostride)[:, :out_n]
        dd = R.DeviceBuffer(nchan * ostride * 4)
        filt.demod_run_batch(din, stride, need, 0.7, dd, ostride, out_n, nchan)
        d = dd.download(np.float32, nchan * ostride).reshape(nchan, ostride)[:, :out_n - 1]
        for c in range(nchan):
            truth = O.fir(xs[c, :n], taps, deci, f64=True)
            e = O.rel_rms(y[c], truth)
            print(f"fir_tcc T={ntaps} D={deci} ch{c} {'tensor' if filt is f else 'fp32'}: {e:.2e}")
            assert e <= 2e-6
            assert O.max_angle_err(d[c] / 0.7, np.angle(truth[1:] * np.conj(truth[:-1]))) <= DEMOD_BAR


def test_fir_tensor_core_translate(R):
    """FirFilter::builder().translate() on the tensor path: pre-rotated (complex) taps + exact-phase rotator epilogue,
    against the FP32 kernels with the same rotator and the reference recurrence (src/fir.rs:416-474); streaming calls keep
    the rotator's output counter."""
    n, deci = 60_000, 2
    x = O.synth_c32(41, 0, n)
    lp = O.low_pass_n(1000.0, 60.0, 101).astype(np.complex64)
    outs = []
    for flags in (0, R.RRC_FIR_NO_TENSOR):
        f = R.Fir(lp, deci=deci, flags=flags)
        f.set_translate(1000.0, 130.0)
        assert f.uses_tensor_cores == (flags == 0)
        first = f.filter(x[:20_000 + 100])              # 10_000 outputs
        rest = f.filter(x[20_000:])                     # the counter carries on
        outs.append(np.concatenate([first, rest]))
    assert len(outs[0]) == len(outs[1])
    assert O.rel_rms(outs[0], outs[1]) <= 2e-6
    rt, ph, st = O.fir_new_translator(lp, 1000.0, 130.0, deci)
    want = O.fir(x, rt, deci)
    O.fir_translate_output(want, ph, st)
    assert O.rel_rms(outs[0][:4000], want[:4000]) < 1e-4      # the reference's f32 rotator recurrence drifts later on


@pytest.mark.parametrize("scale", [1.0, 1e-20, 3e18, 0.0])
def test_fir_tensor_core_block_scaling(R, scale):
    """The per-tile power-of-two scaling makes the fp16 split independent of the stream's level; a tile that is one
    loud burst next to near-silence keeps FP32-class error relative to the burst; all-zero input gives zeros."""
    n, ntaps, deci = 50_000, 129, 4
    taps = O.low_pass_n(1.0, 0.08, ntaps).astype(np.complex64)
    x = (O.synth_c32(31, 0, n) * np.float32(scale)).astype(np.complex64)
    x[20_000:20_050] *= 1000.0
    f = R.Fir(taps, deci=deci)
    assert f.uses_tensor_cores
    y = f.filter(x)
    truth = O.fir(x, taps, deci, f64=True)
    if scale == 0.0:
        assert not y.any()
        return
    e, e_ref = O.rel_rms(y, truth), O.rel_rms(O.fir(x, taps, deci), truth)
    print(f"fir_tc scale {scale:g}: gpu {e:.2e}  f32-oracle {e_ref:.2e}")
    assert e <= 2e-6 and e <= 8 * e_ref


@pytest.mark.parametrize("kind", ["c32_rtaps", "c32_ctaps", "f32"])
@pytest.mark.parametrize("bad", [np.inf, np.nan])
def test_fir_tensor_core_non_finite_samples_stay_local(R, kind, bad):
    """One Inf / NaN sample makes the outputs whose window contains it non-finite (like the reference's sum) — plus, at
    most, the rest of their 8-output block-rows, whose zero-padded Toeplitz taps meet it as 0 * Inf.  Every
    other output of the tile keeps FP32-class accuracy: the tile is still scaled, by its largest FINITE magnitude, even
    though the finite samples would overflow fp16 unscaled."""
    n, ntaps, pos = 30_000, 64, 12_345
    lp = O.low_pass_n(1.0, 0.1, ntaps)
    if kind == "f32":
        taps, x = lp, (O.synth_f32(61, 0, n) * np.float32(3e5)).astype(np.float32)
    else:
        taps = lp.astype(np.complex64) if kind == "c32_rtaps" else (lp * np.exp(0.2j * np.arange(ntaps))).astype(np.complex64)
        x = (O.synth_c32(61, 0, n) * np.float32(3e5)).astype(np.complex64)
    f = R.Fir(taps)
    assert f.uses_tensor_cores
    x[pos] = bad
    y = f.filter(x)
    o = np.arange(len(y))
    touched = (o > pos - ntaps) & (o <= pos)                      # y[o] uses x[o .. o + ntaps)
    far = (o <= pos - 96) | (o > pos + 8)                         # beyond the padded k-range of any block-row that holds `pos`
    assert not np.isfinite(y[touched]).any()
    assert np.isfinite(y[far]).all()
    xz = x.copy(); xz[pos] = 0
    truth = O.fir(xz, taps, 1, f64=True)
    assert O.rel_rms(y[far], truth[far]) <= 2e-6


def test_fir_tensor_core_stopband_dominated_input(R, monkeypatch):
    """A strong out-of-band tone: the error is relative to the INPUT level, so the bar is checked on the hard case."""
    monkeypatch.setenv("RRC_FIR_TENSOR", "2")
    n, ntaps, deci = 200_000, 255, 10
    taps = O.low_pass_n(2.4e6, 100e3, ntaps).astype(np.complex64)
    x = (O.synth_c32(32, 0, n) * 0.05 + 4.0 * np.exp(2j * np.pi * 0.31 * np.arange(n))).astype(np.complex64)
    truth = O.fir(x, taps, deci, f64=True)
    ftc = R.Fir(taps, deci=deci)
    assert ftc.uses_tensor_cores
    e_tc = O.rel_rms(ftc.filter(x), truth)
    e_fp = O.rel_rms(R.Fir(taps, deci=deci, flags=R.RRC_FIR_NO_TENSOR).filter(x), truth)
    e_ref = O.rel_rms(O.fir(x, taps, deci), truth)
    print(f"stopband-dominated: tensor {e_tc:.2e}  fp32 kernel {e_fp:.2e}  f32-oracle {e_ref:.2e}")
    assert e_tc <= 10 * max(e_ref, e_fp)


def test_fir_tensor_core_falls_back(R):
    """translate, complex taps, short or strongly decimating filters and f32 streams stay on the FP32 kernels; u8 I/Q
    input keeps the tensor path (decode fused into its tile load, tests/test_ingest.py)."""
    lp = O.low_pass_n(1.0, 0.1, 64)
    f = R.Fir(lp.astype(np.complex64))
    assert f.uses_tensor_cores
    f.set_input_u8iq(True)
    assert f.uses_tensor_cores
    f.set_input_u8iq(False)
    f.set_translate(1.0, 0.1)
    assert f.uses_tensor_cores                                                                # translate: complex-tap kernel
    f.set_input_u8iq(True)
    assert f.uses_tensor_cores                                                                # ... also from u8 I/Q bytes
    assert R.Fir(lp.astype(np.float32)).uses_tensor_cores                                    # f32 streams: fir_tcf_kernel
    assert not R.Fir(lp.astype(np.float32), deci=3).uses_tensor_cores
    assert R.Fir(cplx_taps(64)).uses_tensor_cores                                             # complex taps: fir_tcc_kernel
    assert not R.Fir(cplx_taps(64), deci=3).uses_tensor_cores
    assert not R.Fir(O.low_pass_n(1.0, 0.1, 15).astype(np.complex64)).uses_tensor_cores
    # the planner keeps decimating short filters (config 3: 255 taps / 10) on the packed-FP32 kernel
    assert not R.Fir(O.low_pass_n(2.4e6, 100e3, 255).astype(np.complex64), deci=10).uses_tensor_cores
    assert R.Fir(O.low_pass_n(1.0, 0.1, 247).astype(np.complex64)).uses_tensor_cores
    assert not R.Fir(O.low_pass_n(1.0, 0.1, 1025).astype(np.complex64)).uses_tensor_cores     # FftFilter territory


def test_fir_huge_deci_falls_back(R):
    """deci so large that the smem tile cannot hold R*deci samples per thread."""
    x = O.synth_c32(14, 0, 400_000)
    taps = cplx_taps(300, 3)
    y = R.Fir(taps, deci=5000).filter(x)
    truth = O.fir(x, taps, 5000, f64=True)
    assert len(y) == len(truth) and O.rel_rms(y, truth) <= REL_RMS_BAR


def test_fir_batch_and_fused_demod(R):
    """rtl_fm shape: nchan channels, 255-tap decimate-by-10 FIR, then QuadratureDemod."""
    nchan, n, ntaps, deci = 7, 30_000, 255, 10
    taps = O.low_pass_n(2.4e6, 100e3, ntaps).astype(np.complex64)
    xs = np.stack([O.synth_c32(100 + c, 0, n) * 0.3 + np.exp(2j * np.pi * 0.01 * (c + 1) * np.arange(n)).astype(np.complex64)
                   for c in range(nchan)])
    f = R.Fir(taps, deci=deci)
    out_n = f.out_count(n)
    need = (out_n - 1) * deci + ntaps
    din = R.DeviceBuffer.from_numpy(xs)
    dy = R.DeviceBuffer(nchan * out_n * 8)
    f.run_batch(din, n, need, dy, out_n, out_n, nchan)
    y = dy.download(np.complex64, nchan * out_n).reshape(nchan, out_n)
    dd = R.DeviceBuffer(nchan * (out_n - 1) * 4)
    f.demod_run_batch(din, n, need, 1.5, dd, out_n - 1, out_n, nchan)
    d = dd.download(np.float32, nchan * (out_n - 1)).reshape(nchan, out_n - 1)
    for c in range(nchan):
        truth = O.fir(xs[c], taps, deci, f64=True)
        assert O.rel_rms(y[c], truth) <= REL_RMS_BAR
        # fused == FirFilter -> QuadratureDemod chain (oracle demod of the f64-truth FIR output)
        want = 1.5 * np.angle(truth[1:] * np.conj(truth[:-1]))
        assert O.max_angle_err(d[c] / 1.5, want / 1.5) <= DEMOD_BAR


def test_fir_translate_matches_reference_kat(R):
    """reference src/fir.rs:744-789 (translate + deci 3 == manual mix then filter)."""
    inp = np.array([complex(i, i * 0.25) for i in range(32)], np.complex64)
    taps = np.array([.5 - .1j, 1 + .2j, -.25 + .05j, .125 - .3j], np.complex64)
    f = R.Fir(taps, deci=3)
    f.set_translate(8.0, 2.0)
    got = f.filter(inp)
    rt, ph, st = O.fir_new_translator(taps, 8.0, 2.0, 3)
    want = O.fir(inp, rt, 3)
    O.fir_translate_output(want, ph, st)
    assert len(got) == len(want) == 9
    assert np.max(np.abs(got - want)) < 1e-3       # the reference's own tolerance
    # longer run: the exact-phase rotator stays within the f32 recurrence's early behaviour
    x = O.synth_c32(15, 0, 4096)
    t2 = O.low_pass_complex(1024.0, 20.0, 10.0)
    f2 = R.Fir(t2)
    f2.set_translate(1024.0, 60.0)
    got2 = f2.filter(x)
    rt2, ph2, st2 = O.fir_new_translator(t2, 1024.0, 60.0, 1)
    want2 = O.fir(x, rt2, 1)
    O.fir_translate_output(want2, ph2, st2)
    assert O.rel_rms(got2, want2) < 1e-4


def test_fir_run_host_count_rule(R):
    x = O.synth_c32(16, 0, 300_000)
    taps = O.low_pass_n(1.0, 0.1, 64).astype(np.complex64)
    y = R.Fir(taps, deci=4).run_host(x)
    truth = O.fir(x, taps, 4, f64=True)
    assert len(y) == len(truth) and O.rel_rms(y, truth) <= REL_RMS_BAR
    assert len(R.Fir(taps, deci=4).run_host(x[:66])) == 0      # < ntaps + deci - 1 -> WaitForStream, nothing out


# ----------------------------------------------------------- FFT filter ---
@pytest.mark.parametrize("ntaps,n", [(1, 5000), (2, 40_000), (193, 8000), (4097, 100_000), (8193, 50_000), (12289, 30_000), (64, 16384 * 3 + 17), (16385, 100_000), (40_000, 90_000)])
def test_fftfilt_matches_f64_convolution(R, ntaps, n):
    taps = (O.low_pass_n(1.0, 0.05, ntaps).astype(np.complex64) * (1 + 0.5j)) if ntaps > 2 else cplx_taps(ntaps)
    x = O.synth_c32(21, 0, n)
    y = R.FftFilt(taps).filter(x)
    truth = O.conv_full_f64_fft(x, taps, n)
    e = O.rel_rms(y, truth)
    S = O.calc_fft_size(ntaps) - ntaps
    e_ref = O.rel_rms(O.fftfilt(x, taps), truth[:(n // S) * S]) if n >= S else float("nan")
    print(f"fftfilt T={ntaps}: gpu {e:.2e}  f32-oracle(overlap-add) {e_ref:.2e}")
    assert e <= REL_RMS_BAR


def test_fftfilt_streaming_state_and_reset(R):
    taps = O.low_pass_n(1.0, 0.1, 301).astype(np.complex64)
    x = O.synth_c32(22, 0, 60_000)
    f = R.FftFilt(taps)
    cuts = [0, 1, 150, 12_345, 12_400, 45_000, 60_000]
    y = np.concatenate([f.filter(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
    truth = O.conv_full_f64_fft(x, taps, len(x))
    assert O.rel_rms(y, truth) <= REL_RMS_BAR
    f.reset()
    assert O.rel_rms(f.filter(x[:5000]), truth[:5000]) <= REL_RMS_BAR


def test_fftfilt_set_history_is_the_left_halo(R):
    """A shard that starts mid-stream gets its ntaps-1 predecessors through set_history."""
    taps = O.low_pass_n(1.0, 0.1, 501).astype(np.complex64)
    x = O.synth_c32(25, 0, 50_000)
    truth = O.conv_full_f64_fft(x, taps, len(x))
    f = R.FftFilt(taps)
    cut = 20_000
    halo = R.DeviceBuffer.from_numpy(x[cut - 500:cut])
    f.set_history(halo, 500)
    assert O.rel_rms(f.filter(x[cut:]), truth[cut:]) <= REL_RMS_BAR
    with pytest.raises(R.RrcError):
        f.set_history(halo, 499)


def test_fftfilt_run_host_reference_count_rule(R):
    """floor(N/nsamples)*nsamples outputs, trailing partial block never flushed (src/fft_filter.rs:315-327)."""
    taps = O.low_pass_complex(8000.0, 1000.0, 100.0)        # 193 taps -> fft 512, block 319
    sig, _ = O.signal_source_complex(8000.0, 3000.0, 1.0, 8000)
    f = R.FftFilt(taps)
    assert (f.ref_fft_size, f.nsamples) == (512, 319)
    y = f.run_host(sig)
    assert len(y) == 7975 == O.fftfilt_out_count(8000, 193)
    # reference test filter_a_signal (src/fft_filter.rs:502-549): stop band < 2e-4 after the transient
    assert np.max(np.abs(y[193:])) < 2e-4
    # The tone sits in the stop band, so |y| ~ 1e-4 |x|: judge the error against the INPUT scale
    # (relative to the almost-cancelled output it would be pure f32 rounding noise of either side).
    truth = O.conv_full_f64(sig, taps, 7975)
    assert np.max(np.abs(y - truth)) <= 1e-5 * np.max(np.abs(sig))
    assert np.max(np.abs(y - truth)) <= 1e-6       # f32 noise floor of a 16384-point transform at unit input


def test_fftfilt_fused_decimation(R):
    taps = O.low_pass_n(1.0, 0.05, 1025).astype(np.complex64)
    x = O.synth_c32(23, 0, 90_000)
    truth = O.conv_full_f64_fft(x, taps, len(x))
    for deci, skip in ((8, 0), (8, 5), (3, 2), (1, 4)):
        f = R.FftFilt(taps)
        din = R.DeviceBuffer.from_numpy(x)
        dout = R.DeviceBuffer(len(x) * 8)
        n = f.decim_run(din, len(x), deci, skip, dout)
        want = truth[skip::deci]
        assert n == len(want)
        assert O.rel_rms(dout.download(np.complex64, n), want) <= REL_RMS_BAR


@pytest.mark.parametrize("ntaps,n", [(1, 5000), (193, 40_000), (4097, 12288 * 5 + 100), (4097, 12288 * 4), (12289, 50_000), (16385, 100_000)])
def test_fftfilt_real_stream_mode(R, ntaps, n):
    """FftFilterFloat compute (src/fft_filter.rs:365-491) with the real-stream kernel mode: two
    consecutive real blocks per complex transform; carried f32 history across calls; host pipeline."""
    taps = O.low_pass_n(1.0, 0.05, ntaps).astype(np.float32) if ntaps > 1 else np.array([0.75], np.float32)
    x = O.synth_f32(33, 0, n)
    truth = O.conv_full_f64_fft(x.astype(np.complex64), taps.astype(np.complex64), n).real
    f = R.FftFilt(taps, real=True)
    got = f.filter(x)
    assert got.dtype == np.float32 and len(got) == n
    assert O.rel_rms(got, truth) <= REL_RMS_BAR
    f.reset()
    cut = n // 3 + 1
    y = np.concatenate([f.filter(x[:cut]), f.filter(x[cut:])])
    assert O.rel_rms(y, truth) <= REL_RMS_BAR
    # equals the reference's construction (widen -> complex filter -> .re) to rounding
    fc = R.FftFilt(taps.astype(np.complex64))
    assert O.rel_rms(got, fc.filter(x.astype(np.complex64)).real) <= 2e-6
    f2 = R.FftFilt(taps, real=True)
    yh = f2.run_host(x)
    assert len(yh) == (n // f2.nsamples) * f2.nsamples
    assert O.rel_rms(yh, truth[:len(yh)]) <= REL_RMS_BAR
    with pytest.raises(R.RrcError):
        f2.decim_run(R.DeviceBuffer(64), 8, 8, 0, R.DeviceBuffer(64))


def test_fftfilt_decim_run_host_chunked(R, monkeypatch):
    """Host pipeline of FftFilter -> RationalResampler(1, deci): chunk boundaries carry the filter
    history and the resampler phase (chunk forced small so several chunks run)."""
    monkeypatch.setenv("RRC_PIPE_CHUNK_LOG2", "12")
    for ntaps, deci, n in ((301, 8, 70_000), (4097, 8, 150_000), (193, 3, 50_001)):
        taps = (O.low_pass_n(1.0, 0.05, ntaps) * (1 - 0.5j)).astype(np.complex64)
        x = O.synth_c32(31, 0, n)
        f = R.FftFilt(taps)
        y = f.decim_run_host(x, deci)
        nfull = (n // f.nsamples) * f.nsamples
        want = O.conv_full_f64_fft(x, taps, n)[:nfull:deci]
        assert len(y) == len(want)
        assert O.rel_rms(y, want) <= REL_RMS_BAR


def test_fftfilt_long_taps_partitioned_streaming_and_decimation(R):
    """BASELINE config 5 shape at test size: 16385 taps (two tap partitions), streamed in pieces, decimate by 8."""
    taps = O.low_pass_n(1.0, 0.02, 16385).astype(np.complex64)
    x = O.synth_c32(24, 0, 150_000)
    truth = O.conv_full_f64_fft(x, taps, len(x))
    f = R.FftFilt(taps)
    y = np.concatenate([f.filter(x[:70_001]), f.filter(x[70_001:])])
    assert O.rel_rms(y, truth) <= REL_RMS_BAR
    f2 = R.FftFilt(taps)
    din = R.DeviceBuffer.from_numpy(x)
    dout = R.DeviceBuffer(len(x) * 8)
    n = f2.decim_run(din, len(x), 8, 0, dout)
    assert n == len(truth[::8])
    assert O.rel_rms(dout.download(np.complex64, n), truth[::8]) <= REL_RMS_BAR


@pytest.mark.parametrize("ntaps,n,skip", [(301, 90_000, 0), (4097, 200_000, 3), (12289, 60_000, 13),      # 16384-point fold kernel
                                          (16385, 400_000, 0), (16385, 70_000, 5), (12290, 300_000, 9),   # 65536-point cluster kernel
                                          (40_001, 250_000, 2), (49_153, 200_000, 7)])
def test_fftfilt_decimate_by_8_folded_spectrum(R, ntaps, n, skip, monkeypatch):
    """FftFilter + RationalResampler(1, 8) fused with the pruned inverse transform (fftfilt_fold.cu):
    against f64 truth, and against the store-predicate path of the plain kernel on the same input."""
    monkeypatch.setenv("RRC_FFTFILT_NO_POLY", "1")              # the polyphase kernel (next test) is the default for deci 8
    taps = (O.low_pass_n(1.0, 0.02, ntaps).astype(np.complex64) * (1 - 0.2j))
    x = O.synth_c32(41, 0, n)
    want = O.conv_full_f64_fft(x, taps, n)[skip::8]
    din = R.DeviceBuffer.from_numpy(x)

    def run():
        f = R.FftFilt(taps)
        dout = R.DeviceBuffer(max(1, len(want)) * 8)
        cnt = f.decim_run(din, n, 8, skip, dout)
        assert cnt == len(want)
        return dout.download(np.complex64, cnt)
    got = run()
    assert O.rel_rms(got, want) <= REL_RMS_BAR
    monkeypatch.setenv("RRC_FFTFILT_NO_FOLD", "1")
    plain = run()
    assert O.rel_rms(plain, want) <= REL_RMS_BAR
    assert O.rel_rms(got, plain) <= REL_RMS_BAR


@pytest.mark.gpu
@pytest.mark.parametrize("C", [0, 4, 2, 1])
@pytest.mark.parametrize("ntaps,n,deci,skip", [(16385, 2_000_000, 8, 0), (16385, 70_000, 8, 5), (301, 90_000, 8, 0), (4097, 200_000, 8, 3),
                                               (5, 20_000, 8, 7), (40_001, 250_000, 8, 2), (4097, 60_000, 3, 1), (64, 50_000, 2, 0),
                                               (16385, 190_000, 16, 9), (12289, 60_000, 5, 13), (1000, 300_000, 4, 2), (5, 5, 8, 0),
                                               (16385, 1_000_001, 8, 20_003)])
def test_fftfilt_polyphase_decimation(R, ntaps, n, deci, skip, C, monkeypatch):
    """FftFilter + RationalResampler(1, deci) as a polyphase filter (fftfilt_poly.cu: deci forward transforms, the sum over the
    branches in tensor memory, one inverse): every cluster width, 128-bit pair gathers and the unaligned fallback, against f64
    truth and against the store-predicate path of the plain kernel."""
    if C and deci % C:
        pytest.skip("cluster width does not divide the decimation")
    monkeypatch.setenv("RRC_FFTFILT_POLY_C", str(C))
    taps = (O.low_pass_n(1.0, 0.02, ntaps).astype(np.complex64) * (1 - 0.2j))
    x = O.synth_c32(43, 0, n)
    want = O.conv_full_f64_fft(x, taps, n)[skip::deci]
    din = R.DeviceBuffer.from_numpy(x)

    def run():
        f = R.FftFilt(taps)
        dout = R.DeviceBuffer(max(1, len(want)) * 8)
        k0 = R.launch_count()
        cnt = f.decim_run(din, n, deci, skip, dout)
        assert cnt == len(want)
        return dout.download(np.complex64, cnt), R.launch_count() - k0
    got, launches = run()
    assert launches == 1                                          # the kernel writes the next history itself
    assert O.rel_rms(got, want) <= REL_RMS_BAR
    monkeypatch.setenv("RRC_FFTFILT_NO_POLY", "1")
    monkeypatch.setenv("RRC_FFTFILT_NO_FOLD", "1")
    plain, _ = run()
    assert O.rel_rms(got, plain) <= REL_RMS_BAR


@pytest.mark.gpu
@pytest.mark.parametrize("u8", [0, 1])
def test_fftfilt_polyphase_streaming_history_epilogue_and_u8(R, u8):
    """Config 5 shape streamed in ragged pieces through the polyphase kernel: the (ntaps-1)-sample history and the decimation
    phase carry across calls; u8 I/Q input (RtlSdrDecode fused into the gather) and a fused ComplexToMag2 store."""
    taps = O.low_pass_n(1.0, 0.02, 16385).astype(np.complex64)
    n = 700_000
    if u8:
        raw = np.frombuffer(np.random.default_rng(5).bytes(2 * n), np.uint8)
        x = O.rtlsdr_decode(raw)
    else:
        x = O.synth_c32(44, 0, n)
    truth = O.conv_full_f64_fft(x, taps, n)[::8]
    for epi in (False, True):
        f = R.FftFilt(taps)
        if u8:
            f.set_input_u8iq(True)
        if epi:
            f.set_epilogue(R.EPI_MAG2)
        pieces, off, got = [300_001, 7, 250_000, 3, 149_989], 0, []
        for m in pieces:
            skip = (-off) % 8
            src = raw[2 * off:2 * (off + m)] if u8 else x[off:off + m]
            din = R.DeviceBuffer.from_numpy(np.ascontiguousarray(src))
            dout = R.DeviceBuffer(max(1, m // 8 + 2) * 8)
            cnt = f.decim_run(din, m, 8, skip, dout)
            got.append(dout.download(np.float32 if epi else np.complex64, cnt))
            off += m
        got = np.concatenate(got)
        want = (truth.real.astype(np.float64) ** 2 + truth.imag.astype(np.float64) ** 2) if epi else truth
        assert len(got) == len(want)
        assert O.rel_rms(got, want) <= (3e-5 if epi else REL_RMS_BAR)


def test_fftfilt_fold_streaming_carries_history_and_phase(R):
    """Config 5 shape streamed in ragged pieces: the (ntaps-1)-sample history and the decimation phase
    (RationalResampler's counter, src/rational_resampler.rs:101-105) carry across calls."""
    taps = O.low_pass_n(1.0, 0.02, 16385).astype(np.complex64)
    x = O.synth_c32(42, 0, 330_000)
    truth = O.conv_full_f64_fft(x, taps, len(x))[::8]
    f = R.FftFilt(taps)
    out, pos = [], 0
    for cut in (5, 70_001, 70_004, 200_000, 330_000):
        piece = x[pos:cut]
        skip = (-pos) % 8
        din = R.DeviceBuffer.from_numpy(piece)
        dout = R.DeviceBuffer(max(1, len(piece)) * 8)
        cnt = f.decim_run(din, len(piece), 8, skip, dout)
        out.append(dout.download(np.complex64, cnt))
        pos = cut
    y = np.concatenate(out)
    assert len(y) == len(truth)
    assert O.rel_rms(y, truth) <= REL_RMS_BAR


# ------------------------------------------------------------ resampler ---
@pytest.mark.parametrize("interp,deci", [(1, 1), (1, 2), (2, 1), (2, 3), (3, 2), (25, 64), (25, 128), (147, 160),
                                          (200000, 1024000), (1, 8), (160, 147), (7, 1000), (1000, 7)])
@pytest.mark.parametrize("dtype", [np.uint8, np.uint32, np.complex64])
def test_resampler_bit_exact(R, interp, deci, dtype):
    n = 100_003
    if dtype == np.complex64:
        x = O.synth_c32(31, 0, n)
    else:
        x = (np.arange(n) * 2654435761 % (np.iinfo(dtype).max + 1)).astype(dtype)
    want = O.resample(x, interp, deci)
    r = R.Resampler(x.dtype.itemsize, interp, deci)
    w, consumed, got = r.work(x, len(want) + 10)
    assert w == 0 and consumed == n                     # WaitForStream(src, 1)
    assert got.tobytes() == want.tobytes()
    assert len(got) == O.resample_out_count(n, interp, deci)


def test_resampler_kat_example64(R):
    """reference src/rational_resampler.rs:248-262"""
    r = R.Resampler(4, 25, 64)
    _, _, got = r.work(np.arange(50, dtype=np.uint32), 100)
    assert list(got) == [0, 2, 5, 7, 10, 12, 15, 17, 20, 23, 25, 28, 30, 33, 35, 38, 40, 43, 46, 48]


def test_resampler_work_sequence_matches_oracle_state(R):
    """Random input/output window sizes: consumed/produced/wait/counter/pending equal the restated work()."""
    rng = np.random.default_rng(5)
    for interp, deci in ((3, 1), (7, 3), (147, 160), (1, 9), (11, 2)):
        x = (np.arange(40_000) * 7919 % 65521).astype(np.uint32)
        g, o = R.Resampler(4, interp, deci), O.Resampler(4, interp, deci)
        pos = 0
        outs_g, outs_o = [], []
        for _ in range(400):
            n_in = int(rng.integers(0, 300))
            cap = int(rng.integers(0, 200))
            win = x[pos:pos + n_in]
            ro, co, yo = o.work(win, cap)
            wg, cg, yg = g.work(win, cap)
            assert (cg, len(yg)) == (co, len(yo))
            assert yg.tobytes() == yo.tobytes()
            assert wg == (1 if ro == O.Resampler.WAIT_DST else 0)
            gi, gd, gc, gp = g.state()
            assert gp == o.has_pending
            assert gc == o.counter
            pos += cg
            outs_g.append(yg)
        assert pos > 0


def test_resampler_interpolation_survives_full_output_buffer(R):
    """reference src/rational_resampler.rs:278-299"""
    cap = 4_096_000 // 4
    boundary = cap // 3
    x = np.arange(boundary + 1, dtype=np.uint32)
    r = R.Resampler(4, 3, 1)
    w, consumed, first = r.work(x, cap)
    assert w == 1 and len(first) == cap and first[-1] == boundary and consumed == boundary + 1
    assert r.state()[3]                                  # pending sample
    w, consumed, second = r.work(np.empty(0, np.uint32), cap)
    assert w == 0 and list(second) == [boundary, boundary] and not r.state()[3]


def test_resampler_zero_is_error(R):
    for i, d in ((0, 1), (1, 0)):
        with pytest.raises(R.RrcError):
            R.Resampler(4, i, d)


def test_resampler_run_host(R):
    x = O.synth_f32(33, 0, 1_000_000)
    want = O.resample(x, 147, 160)
    r = R.Resampler(4, 147, 160)
    consumed, got = r.run_host(x, len(want) + 5)
    assert consumed == len(x) and got.tobytes() == want.tobytes()


# ---------------------------------------------------------------- demod ---
def test_quad_demod(R):
    n = 200_000
    x = (np.exp(2j * np.pi * np.cumsum(0.05 * np.sin(2 * np.pi * 1e-3 * np.arange(n)))) * (1 + 0.1 * O.synth_f32(41, 0, n))).astype(np.complex64)
    got = R.quad_demod_host(x, 0.7)
    want = O.quad_demod(x, 0.7, f64=True)
    assert len(got) == n - 1
    assert O.max_angle_err(got / 0.7, want / 0.7) <= DEMOD_BAR
    print("demod max err rad:", O.max_angle_err(got / 0.7, want / 0.7), " f32 oracle:", O.max_angle_err(O.quad_demod(x, 0.7) / 0.7, want / 0.7))
    # reference KATs src/quadrature_demod.rs:211-264
    assert list(R.quad_demod_host(np.zeros(4, np.complex64))) == [0.0, 0.0, 0.0]
    cw = R.quad_demod_host(np.array([1, 0.707 - 0.707j, -1j, -1], np.complex64))
    assert np.allclose(cw, [-np.pi / 4, -np.pi / 4, -np.pi / 2], atol=1e-3)
    # white noise input: angles all over (-pi, pi]
    z = O.synth_c32(42, 0, 100_000)
    assert O.max_angle_err(R.quad_demod_host(z), O.quad_demod(z, f64=True)) <= DEMOD_BAR


# ----------------------------------------- BASELINE-size spot checks -------
def _spot_check_fir(R, x_dev, y_dev, n_out, taps, deci, seed, n_in, npts=400):
    """Random outputs of a device-resident run against f64 dot products of the
    regenerated synthetic input (the input never has to exist on the host)."""
    rng = np.random.default_rng(0)
    idx = np.unique(np.concatenate([[0, 1, n_out - 1], rng.integers(0, n_out, npts)]))
    T = len(taps)
    got = np.array([y_dev.download(np.complex64, 1, int(i) * 8)[0] for i in idx])
    want = np.empty(len(idx), np.complex128)
    rev = taps[::-1].astype(np.complex128)
    for k, i in enumerate(idx):
        w = O.synth_c32(seed, int(i) * deci, T).astype(np.complex128)
        want[k] = np.dot(w, rev)
    return O.rel_rms(got, want)


def test_fir_config1_full_size(R):
    """BASELINE config 1: 64-tap c32 low-pass, no decimation, 2^24 samples."""
    n, seed = 1 << 24, 0x5EED0001
    taps = O.low_pass_n(1.0, 0.1, 64).astype(np.complex64)
    f = R.Fir(taps)
    din = R.DeviceBuffer(n * 8)
    R.synth_f32(din, seed, 0, 2 * n)
    n_out = f.out_count(n)
    assert n_out == 16_777_153
    dout = R.DeviceBuffer(n_out * 8)
    f.run(din, n, dout, n_out)
    assert f.kernel_name.startswith("fir_tc5_kernel")          # 2048 tiles >= 3 per SM: the tcgen05 kernel by default
    assert _spot_check_fir(R, din, dout, n_out, taps, 1, seed, n) <= REL_RMS_BAR
    small = R.Fir(taps)
    small.filter(O.synth_c32(3, 0, 100_000))
    assert small.kernel_name.startswith("fir_tc1_kernel")      # 13 tiles: the mma.sync walk kernel
    # the device generator and the oracle generator are the same function
    h = din.download(np.complex64, 4096)
    assert np.array_equal(h, O.synth_c32(seed, 0, 4096))


def test_fftfilt_config2_full_size(R):
    """BASELINE config 2: 4097-tap FftFilter over 2^28 c32 samples (device-resident)."""
    n, seed, T = 1 << 28, 0x5EED0002, 4097
    taps = O.low_pass_n(1.0, 0.05, T).astype(np.complex64)
    f = R.FftFilt(taps)
    din = R.DeviceBuffer(n * 8)
    R.synth_f32(din, seed, 0, 2 * n)
    n_out = O.fftfilt_out_count(n, T)
    assert n_out == 268_434_089
    dout = R.DeviceBuffer(n * 8)
    f.run(din, n_out, dout)
    rng = np.random.default_rng(1)
    idx = np.unique(np.concatenate([[0, 1, T - 2, T - 1, T, 12287, 12288, n_out - 1], rng.integers(0, n_out, 300)]))
    got = np.array([dout.download(np.complex64, 1, int(i) * 8)[0] for i in idx])
    want = np.empty(len(idx), np.complex128)
    h64 = taps.astype(np.complex128)
    for k, i in enumerate(idx):
        i = int(i)
        lo = max(0, i - T + 1)
        w = O.synth_c32(seed, lo, i - lo + 1).astype(np.complex128)     # x[lo..i]
        want[k] = np.dot(w[::-1], h64[:len(w)])                          # sum_k h[k] x[i-k]
    assert O.rel_rms(got, want) <= REL_RMS_BAR


# ------------------------------------------------ time-segment sharding ------
def test_time_segment_shards_equal_whole(R):
    """splitting a stream by time segment with an ntaps-1 halo (0 for the resampler) and
    running each shard through the CUDA kernels independently reproduces the whole-stream result."""
    from rustradio_b200 import shard as S
    n, world = 300_000, 4
    x = O.synth_c32(51, 0, n)
    taps = O.low_pass_n(1.0, 0.08, 257).astype(np.complex64)
    whole_fir = O.fir(x, taps, 5, f64=True)
    parts = []
    for r in range(world):
        seg = S.fir_segment(n, len(taps), 5, world, r)
        y = R.Fir(taps, deci=5).filter(x[seg.in_lo:seg.in_hi])
        assert len(y) == seg.out_hi - seg.out_lo
        parts.append(y)
    assert O.rel_rms(np.concatenate(parts), whole_fir) <= REL_RMS_BAR
    n_out = O.fftfilt_out_count(n, len(taps))
    whole_fft = O.conv_full_f64_fft(x, taps, n_out)
    parts = []
    for r in range(world):
        seg = S.fftfilt_segment(n, len(taps), world, r)
        y = R.FftFilt(taps).filter(x[seg.in_lo:seg.in_hi])       # fresh filter: zero history before the halo
        parts.append(y[seg.out_lo - seg.in_lo:])
    got = np.concatenate(parts)
    assert len(got) == n_out and O.rel_rms(got, whole_fft) <= REL_RMS_BAR
    xf = O.synth_f32(52, 0, n)
    whole_rs = O.resample(xf, 147, 160)
    parts = []
    for r in range(world):
        seg = S.resampler_segment(n, 147, 160, world, r)
        # a shard starts mid-stream: rrc_resampler_set_state seeds the reference's `counter`
        # (src/rational_resampler.rs:101-105) so that the shard's first output is global output out_lo;
        # EVERY shard runs through resample_kernel
        res = R.Resampler(4, 147, 160)
        c0 = seg.in_lo * 147 - seg.out_lo * 160
        assert -147 < c0 <= 0
        res.set_state(c0)
        _, consumed, y = res.work(xf[seg.in_lo:seg.in_hi], seg.out_hi - seg.out_lo)
        assert len(y) == seg.out_hi - seg.out_lo
        parts.append(y)
    assert np.concatenate(parts).tobytes() == whole_rs.tobytes()


def test_fftfilt_halo_through_history_pointer(R):
    """SURVEY 8(e) time-segment sharding, the way bench.py does it across GPUs: shard r > 0 reads its
    ntaps-1 sample left halo through rrc_fftfilt_set_history_ptr (zero-copy: the pointer is the tail of
    the left neighbour's input buffer) or has it copied by rrc_fftfilt_set_history; both equal the whole
    stream.  Also covers the fused decimate-by-8 (fold) kernel and that the pointer is one-shot."""
    from rustradio_b200 import shard as S
    n, world = 400_000, 3
    x = O.synth_c32(53, 0, n)
    for T, deci in ((257, 1), (4097, 1), (4097, 8)):
        taps = O.low_pass_n(1.0, 0.08, T).astype(np.complex64)
        n_out = O.fftfilt_out_count(n, T)
        whole = O.conv_full_f64_fft(x, taps, n_out)[::deci]
        dx = R.DeviceBuffer.from_numpy(x)                       # the "neighbour's" buffer: all shards' inputs
        for mode in ("ptr", "copy"):
            parts = []
            for r in range(world):
                seg = S.fftfilt_segment(n, T, world, r)
                m = seg.out_hi - seg.out_lo
                f = R.FftFilt(taps)
                if r > 0:
                    halo = dx.ptr + (seg.out_lo - (T - 1)) * 8
                    f.set_history_ptr(halo, T - 1) if mode == "ptr" else f.set_history(halo, T - 1)
                skip = (-seg.out_lo) % deci
                cnt = (m - skip + deci - 1) // deci if m > skip else 0
                dout = R.DeviceBuffer(max(cnt, 1) * 8)
                if deci == 1:
                    f.run(dx.ptr + seg.out_lo * 8, m, dout)
                else:
                    assert f.decim_run(dx.ptr + seg.out_lo * 8, m, deci, skip, dout) == cnt
                parts.append(dout.download(np.complex64, cnt))
                if r == 1 and deci == 1:                        # one-shot: a second run continues from the carried history
                    more = min(1000, n - seg.out_hi)
                    d2 = R.DeviceBuffer(more * 8)
                    f.run(dx.ptr + seg.out_hi * 8, more, d2)
                    ref = O.conv_full_f64_fft(x, taps, seg.out_hi + more)[seg.out_hi:]
                    assert O.rel_rms(d2.download(np.complex64, more), ref) <= REL_RMS_BAR
            got = np.concatenate(parts)
            assert len(got) == len(whole) and O.rel_rms(got, whole) <= REL_RMS_BAR, (T, deci, mode)


# ------------------------------------------------------ empty / ragged inputs ---
def test_empty_and_ragged_inputs(R):
    """Zero-length and shorter-than-one-output inputs are no-ops with the reference's counts."""
    taps = O.low_pass_n(1.0, 0.1, 33).astype(np.complex64)
    f = R.Fir(taps, deci=4)
    assert len(f.filter(np.empty(0, np.complex64))) == 0
    assert len(f.filter(O.synth_c32(1, 0, 35))) == 0            # < ntaps + deci - 1 = 36
    assert len(f.filter(O.synth_c32(1, 0, 36))) == 1
    assert len(f.run_host(np.empty(0, np.complex64))) == 0
    g = R.FftFilt(taps)
    assert len(g.filter(np.empty(0, np.complex64))) == 0
    assert len(g.run_host(O.synth_c32(1, 0, g.nsamples - 1))) == 0    # partial block is never flushed
    assert len(g.run_host(O.synth_c32(1, 0, g.nsamples))) == g.nsamples
    r = R.Resampler(8, 3, 7)
    w, c, y = r.work(np.empty(0, np.complex64), 10)
    assert (w, c, len(y)) == (0, 0, 0)                                # WaitForStream(src, 1)
    w, c, y = r.work(O.synth_c32(1, 0, 5), 0)
    assert (w, c, len(y)) == (1, 0, 0)                                # WaitForStream(dst, 1)
    assert len(R.quad_demod_host(np.empty(0, np.complex64))) == 0
    assert len(R.quad_demod_host(np.ones(1, np.complex64))) == 0
    assert len(R.quad_demod_host(np.ones(2, np.complex64))) == 1
