"""RtlSdrDecode: oracle pinned on the reference's golden vector, the integer
part of work(), the block model, and — on the GPU — the decode kernel, the decode fused into the
FIR / FftFilter first load, and the rr::RtlSdrDecode block driven like the reference's #[test]s
(src/rtlsdr_decode.rs:51-137)."""
import numpy as np
import pytest

from oracle import blockmodel as B
from oracle import oracle as O

GOLD_IN = [0, 10, 20, 10, 0, 13]
# src/rtlsdr_decode.rs:75-92 (assert_eq!, exact)
GOLD_OUT = np.array([-1.016, -0.93600005, -0.85600007, -0.93600005, -1.016, -0.91200006], np.float32)


# ------------------------------------------------------------------ CPU ---
def test_oracle_matches_reference_golden_vector():
    got = O.rtlsdr_decode(GOLD_IN)
    assert got.view(np.float32).tobytes() == GOLD_OUT.tobytes()


def test_oracle_all_byte_values_and_odd_length():
    raw = np.arange(256, dtype=np.uint8).repeat(2)
    got = O.rtlsdr_decode(raw)
    want = ((np.arange(256, dtype=np.float32) - np.float32(127.0)) * np.float32(0.008)).astype(np.float32)
    assert got.real.tobytes() == want.tobytes() and got.imag.tobytes() == want.tobytes()
    assert len(O.rtlsdr_decode([0, 10, 20, 10, 0])) == 2          # `uneven`, :97-106
    assert len(O.rtlsdr_decode([])) == 0                          # `empty`, :58-68


def test_blockmodel_overflow_sequence():
    """`overflow` (:108-137): DEFAULT_STREAM_SIZE input bytes come out in 4 rounds of SIZE/8 samples."""
    size = B.DEFAULT_STREAM_SIZE
    src = B.Stream(np.uint8)
    src.write_buf()[:size] = 0
    src.produce(size, [])
    dec = B.RtlSdrDecode(src)
    for _ in range(4):
        assert dec.work().kind == B.WAIT
        res, _t = dec.out.read_buf()
        assert len(res) == size // 8
        dec.out.consume(size // 8)
    r = dec.work()
    assert r.kind == B.WAIT and r.stream is src and r.need == 2
    assert len(dec.out.read_buf()[0]) == 0


def test_plan_matches_blockmodel():
    import rustradio_b200 as R
    cap = B.DEFAULT_STREAM_SIZE // 8
    for in_len, free in [(0, 10), (1, 10), (2, 0), (2, 1), (5, 10), (6, 2), (7, 2), (4_096_000, cap), (999, 10_000)]:
        src = B.Stream(np.uint8)
        src.write_buf()[:in_len] = 1
        src.produce(in_len, [])
        dec = B.RtlSdrDecode(src)
        used = cap - free                                         # fill the output so that `free` remain
        dec.out.produce(used, [])
        r = dec.work()
        consume, produce, need, on_out = R.rtlsdr_decode_plan(in_len, free)
        assert produce == len(dec.out.read_buf()[0]) - used
        assert consume == in_len - len(src.read_buf()[0])
        assert (need, bool(on_out)) == (r.need, r.stream is dec.out)


# ------------------------------------------------------------------ GPU ---
@pytest.fixture(scope="module")
def R():
    import rustradio_b200 as R
    assert R.device_count() >= 1
    return R


@pytest.mark.gpu
@pytest.mark.parametrize("n_bytes,in_off,out_off", [(6, 0, 0), (0, 0, 0), (1, 0, 0), (5, 0, 0), (2_000_001, 0, 0),
                                                    (100_000, 2, 0), (100_000, 6, 1), (100_001, 3, 0), (4096, 14, 1)])
def test_decode_kernel_bit_exact(R, n_bytes, in_off, out_off):
    raw = np.array(GOLD_IN, np.uint8) if n_bytes == 6 else O.synth_u8(11, 0, n_bytes)
    din = R.DeviceBuffer.from_numpy(np.concatenate([np.zeros(in_off, np.uint8), raw])) if n_bytes + in_off else R.DeviceBuffer(16)
    dout = R.DeviceBuffer((n_bytes // 2 + out_off + 1) * 8)
    R.rtlsdr_decode(din.ptr + in_off, n_bytes, dout.ptr + 8 * out_off)
    got = dout.download(np.complex64, n_bytes // 2 + out_off)[out_off:]
    assert got.tobytes() == O.rtlsdr_decode(raw).tobytes()
    if n_bytes == 6:
        assert got.view(np.float32).tobytes() == GOLD_OUT.tobytes()
    assert R.rtlsdr_decode_host(raw).tobytes() == O.rtlsdr_decode(raw).tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("ntaps,deci,n,kind", [(64, 1, 50_000, "rtaps"), (255, 10, 240_000, "rtaps"), (33, 3, 20_001, "ctaps"),
                                               (7, 1, 4000, "rtaps"), (300, 1000, 10_000, "rtaps"), (64, 1, 50_001, "rtaps"),
                                               (200, 2, 30_000, "rtaps"), (121, 1, 777, "rtaps"), (160, 4, 44_444, "rtaps"), (64, 1, 50_001, "ctaps"), (200, 2, 30_000, "ctaps")])
def test_fir_fused_u8_input_equals_decode_then_fir(R, ntaps, deci, n, kind):
    """RtlSdrDecode -> FirFilter<Complex> fused: bit-identical to decoding first, and within the
    FIR bar of the f64 truth."""
    raw = O.synth_u8(12, 0, 2 * n)
    x = O.rtlsdr_decode(raw)
    taps = O.low_pass_n(1.0, 0.1, ntaps).astype(np.complex64)
    if kind == "ctaps":
        taps = (taps * np.exp(1j * 0.3 * np.arange(ntaps))).astype(np.complex64)
    f = R.Fir(taps, deci=deci)
    want = f.filter(x)
    f.set_input_u8iq(True)
    n_out = f.out_count(n)
    need = (n_out - 1) * deci + ntaps
    din = R.DeviceBuffer.from_numpy(raw)
    dout = R.DeviceBuffer(max(n_out, 1) * 8)
    f.run(din, need, dout, n_out)
    got = dout.download(np.complex64, n_out)
    assert got.tobytes() == want.tobytes()
    assert O.rel_rms(got, O.fir(x, taps, deci, f64=True)) <= 1e-5
    assert f.run_host(raw).tobytes() == want.tobytes()


@pytest.mark.gpu
def test_fir_demod_fused_u8_rtl_fm_chain(R):
    """The rtl_fm chain RtlSdrDecode -> FirFilter(255 taps, /10) -> QuadratureDemod in ONE kernel."""
    nchan, n = 3, 60_000
    raw = O.synth_u8(13, 0, 2 * n * nchan)
    taps = O.low_pass_n(2.4e6, 100e3, 255).astype(np.complex64)
    f = R.Fir(taps, deci=10)
    n_out = f.out_count(n)
    need = (n_out - 1) * 10 + 255
    dd = R.DeviceBuffer(nchan * (n_out - 1) * 4)
    x = O.rtlsdr_decode(raw)
    f.demod_run_batch(R.DeviceBuffer.from_numpy(x), n, need, 0.5, dd, n_out - 1, n_out, nchan)
    want = dd.download(np.float32, nchan * (n_out - 1))
    f.set_input_u8iq(True)
    dd2 = R.DeviceBuffer(nchan * (n_out - 1) * 4)
    f.demod_run_batch(R.DeviceBuffer.from_numpy(raw), n, need, 0.5, dd2, n_out - 1, n_out, nchan)
    got = dd2.download(np.float32, nchan * (n_out - 1))
    assert got.tobytes() == want.tobytes()
    for c in range(nchan):
        fy = O.fir(x[c * n:(c + 1) * n], taps, 10, f64=True)
        assert O.max_angle_err(got[c * (n_out - 1):(c + 1) * (n_out - 1)] / 0.5, np.angle(fy[1:] * np.conj(fy[:-1]))) <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("ntaps,n", [(4097, 100_000), (193, 40_000), (16385, 120_000)])
def test_fftfilt_fused_u8_input(R, ntaps, n):
    raw = O.synth_u8(14, 0, 2 * n)
    x = O.rtlsdr_decode(raw)
    taps = (O.low_pass_n(1.0, 0.05, ntaps) * (1 + 0.25j)).astype(np.complex64)
    want = R.FftFilt(taps).filter(x)
    f = R.FftFilt(taps)
    f.set_input_u8iq(True)
    cut = 2 * (n // 3)                                             # two calls: history is carried as c32
    d1, d2 = R.DeviceBuffer.from_numpy(raw[:cut]), R.DeviceBuffer.from_numpy(raw[cut:])
    o1, o2 = R.DeviceBuffer(cut // 2 * 8), R.DeviceBuffer((n - cut // 2) * 8)
    f.run(d1, cut // 2, o1)
    f.run(d2, n - cut // 2, o2)
    got = np.concatenate([o1.download(np.complex64, cut // 2), o2.download(np.complex64, n - cut // 2)])
    assert O.rel_rms(got, O.conv_full_f64_fft(x, taps, n)) <= 1e-5
    assert O.rel_rms(got, want) <= 2e-6
    # fused decimate-by-8 (config 5 shape) and the host pipeline
    f2 = R.FftFilt(taps)
    f2.set_input_u8iq(True)
    dd = R.DeviceBuffer((n // 8 + 8) * 8)
    cnt = f2.decim_run(R.DeviceBuffer.from_numpy(raw), n, 8, 3, dd)
    truth = O.conv_full_f64_fft(x, taps, n)[3::8]
    assert cnt == len(truth)
    assert O.rel_rms(dd.download(np.complex64, cnt), truth) <= 1e-5
    f4 = R.FftFilt(taps)
    f4.set_input_u8iq(True)
    yd = f4.decim_run_host(raw, 8)
    nfull = (n // f4.nsamples) * f4.nsamples
    assert len(yd) == (nfull + 7) // 8
    assert O.rel_rms(yd, O.conv_full_f64_fft(x, taps, n)[:nfull:8]) <= 1e-5
    f3 = R.FftFilt(taps)
    f3.set_input_u8iq(True)
    yh = f3.run_host(raw)
    assert len(yh) == (n // f3.nsamples) * f3.nsamples
    assert O.rel_rms(yh, O.conv_full_f64_fft(x, taps, n)[:len(yh)]) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("res", ["DEVICE", "HOST"])
def test_block_reference_tests(R, res):
    """src/rtlsdr_decode.rs `empty`, `some_input`, `uneven`, `overflow`."""
    from rustradio_b200 import blocks as K
    r = getattr(K, res)
    # some_input / uneven
    for data, want_n in (([0, 10, 20, 10, 0, 13], 3), ([0, 10, 20, 10, 0], 2)):
        src, so = K.VectorSource(np.array(data, np.uint8), residency=r)
        assert src.work().kind == K.EOF
        dec, do = K.RtlSdrDecode(so, residency=r)
        ret = dec.work()
        assert ret.kind == K.WAIT
        got, tags = do.read_buf()
        assert len(got) == want_n and tags == []
        assert got.view(np.float32).tobytes() == GOLD_OUT[:2 * want_n].tobytes()
    # overflow: 4 rounds of SIZE/8 samples, then WaitForStream(src, 2)
    size = K.DEFAULT_STREAM_SIZE
    src, so = K.VectorSource(np.zeros(size, np.uint8), residency=r)
    assert src.work().kind == K.EOF
    dec, do = K.RtlSdrDecode(so, residency=r)
    # Host rings have the reference's exact capacity (SIZE/8 samples); device rings round the
    # capacity up to the 2 MiB VMM granularity, so the rounds are larger and fewer.
    left, rounds = size // 2, 0
    while left:
        assert dec.work().kind == K.WAIT
        got, _ = do.read_buf()
        if res == "HOST":
            assert len(got) == size // 8
        else:
            assert len(got) == min(left, 4 * 1024 * 1024 // 8)
        assert np.all(got == np.complex64(complex(-1.016, -1.016)))
        do.consume(len(got))
        left -= len(got)
        rounds += 1
    assert rounds == (4 if res == "HOST" else 4)
    assert dec.work().kind == K.WAIT
    assert len(do.read_buf()[0]) == 0


@pytest.mark.gpu
def test_graph_rtlsdr_decode_fir_demod_chain(R):
    """examples/rtl_fm.rs shape: VectorSource<u8> -> RtlSdrDecode -> FirFilter(/10) -> QuadratureDemod."""
    from rustradio_b200 import blocks as K
    n = 300_000
    raw = O.synth_u8(15, 0, 2 * n)
    taps = O.low_pass_n(2.4e6, 100e3, 255).astype(np.complex64)
    src, s0 = K.VectorSource(raw)
    dec, s1 = K.RtlSdrDecode(s0)
    fir, s2 = K.FirFilter(s1, taps, 10)
    dem, s3 = K.QuadratureDemod(s2, 1.0)
    K.graph_run([src, dec, fir, dem])
    got, _ = s3.read_buf()
    x = O.rtlsdr_decode(raw)
    fy = O.fir(x, taps, 10, f64=True)
    assert len(got) == len(fy) - 1
    assert O.max_angle_err(got, np.angle(fy[1:] * np.conj(fy[:-1]))) <= 1e-4
