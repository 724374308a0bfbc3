"""SURVEY 8f ranks 3a and 4 — Hilbert and the sample-wise `sync` blocks next to the filters
(MultiplyConst, AddConst, ComplexToMag2, Tee, IqBalance).

CPU: the oracle against the reference's own tests for these blocks (src/window.rs doc-test and
`one_tap_windows_are_unity`, src/tee.rs `simple`, src/iq_balance.rs
`removes_dc_offset_quickly_with_large_alpha`; src/hilbert.rs has no #[test], so Hilbert is pinned
through its defining properties) and the block models; the library's host-side tap design against the
oracle bit for bit.  GPU: kernels and rr:: blocks against the oracle through the C ABI."""
import numpy as np
import pytest

from oracle import blockmodel as B
from oracle import oracle as O


def fill(stream, data, tags=()):
    w = stream.write_buf()
    w[:len(data)] = data
    stream.produce(len(data), list(tags))


# ------------------------------------------------------------------ CPU ---
def test_window_reference_tests():
    # src/window.rs doc-test (Hamming, 3 taps, tolerance 0.1) and one_tap_windows_are_unity
    w = O.make_window(O.WINDOW_HAMMING, 3)
    assert np.all(np.abs(w - np.array([0.0869565, 1.0, 0.0869565], np.float32)) < 0.1)
    for wt in (O.WINDOW_HAMMING, O.WINDOW_BLACKMAN, O.WINDOW_BLACKMAN_HARRIS):
        assert O.make_window(wt, 1).tolist() == [1.0]


def test_library_tap_design_equals_oracle():
    import rustradio_b200 as R
    for wt in range(3):
        for n in (1, 2, 3, 5, 65, 129, 1001):
            w = O.make_window(wt, n)
            assert R.make_window(wt, n).tobytes() == w.tobytes()
            if n >= 2:
                assert R.hilbert_taps(w).tobytes() == O.hilbert_taps(w).tobytes()
    assert R.make_window(R.WINDOW_HAMMING_PARM, 65, 0.54).tobytes() == O.make_window(3, 65, 0.54).tobytes()
    for fs, tau in [(2_400_000, 0.2), (48_000, 0.5), (0, 0.2), (1000, float("nan")), (1000, -1.0)]:
        assert R.iq_balance_alpha_from_tau(fs, tau) == O.iq_balance_alpha_from_tau(fs, tau)
    with pytest.raises(R.RrcError):
        R.hilbert_taps(np.ones(1, np.float32))
    with pytest.raises(R.RrcError):
        R.make_window(7, 5)


def test_hilbert_taps_structure():
    """fir::hilbert (src/fir.rs:660-680): odd-symmetric, zero at the centre and at even offsets."""
    for n in (3, 65, 127):
        t = O.hilbert_taps(O.make_window(O.WINDOW_HAMMING, n))
        mid = (n - 1) // 2
        assert t[mid] == 0.0
        even = [mid + i for i in range(-mid, mid + 1) if i % 2 == 0]
        assert np.all(t[even] == 0.0) and np.all(t[[mid - 1, mid + 1]] != 0.0)
        assert np.allclose(t[mid + 1:], -t[:mid][::-1], atol=1e-7)


def test_hilbert_oracle_properties():
    """out.re is the input delayed by (ntaps+1)/2; a cosine comes out as an analytic signal; the
    result does not depend on how work() calls slice the stream (history carry, src/hilbert.rs:123)."""
    T = 65
    t = np.arange(6000)
    x = np.cos(2 * np.pi * 0.11 * t).astype(np.float32)
    y = O.Hilbert(T).work(x)
    d = (T + 1) // 2
    assert np.array_equal(y.real[d:], x[:-d]) and np.all(y.real[:d] == 0)
    assert np.all(np.abs(np.abs(y[2 * T:]) - 1.0) < 1e-2)             # 65-tap Hamming design: < 1 % ripple
    # upper sideband: the phase advances by +2*pi*0.11 per sample
    dphi = np.angle(y[2 * T + 1:] * np.conj(y[2 * T:-1]))
    assert np.all(np.abs(dphi - 2 * np.pi * 0.11) < 1e-2)
    h = O.Hilbert(T)
    cuts = [0, 1, 2, 64, 65, 66, 700, 701, 3000, 6000]
    pieces = [h.work(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.concatenate(pieces).tobytes() == y.tobytes()
    assert O.rel_rms(y, O.Hilbert(T).work(x, f64=True)) < 1e-6


def test_hilbert_blockmodel_counts_and_tags():
    src = B.Stream(np.float32)
    x = O.synth_f32(31, 0, 1000)
    fill(src, x, [B.tag_bool(0, "a", True), B.tag_u64(999, "b", 7)])
    h = B.Hilbert(src, 65)
    assert h.work().kind == B.AGAIN
    r = h.work()
    assert r.kind == B.WAIT and r.stream is src and r.need == 1
    got, tags = h.out.read_buf()
    assert len(got) == 1000
    assert tags == [B.tag_bool(0, "a", True), B.tag_u64(999, "b", 7)]
    assert got.tobytes() == O.Hilbert(65).work(x).tobytes()
    with pytest.raises(AssertionError):
        B.Hilbert(B.Stream(np.float32), 64)


def test_sync_oracles_bit_patterns():
    x = O.synth_c32(32, 0, 1000)
    v = np.complex64(0.3 - 1.7j)
    # num-complex order: (ac - bd) + (ad + bc)i with every operation rounded to f32
    a, b, c, d = x.real, x.imag, np.float32(v.real), np.float32(v.imag)
    want = (a * c - b * d) + 1j * (a * d + b * c)
    assert O.multiply_const(x, v).tobytes() == want.astype(np.complex64).tobytes()
    assert O.add_const(x, v).tobytes() == (x + v).astype(np.complex64).tobytes()
    assert O.complex_to_mag2(x).tobytes() == (a * a + b * b).astype(np.float32).tobytes()
    f = O.synth_f32(33, 0, 1000)
    assert O.multiply_const(f, 0.37).tobytes() == (f * np.float32(0.37)).tobytes()
    assert O.add_const(f, -2.5).tobytes() == (f + np.float32(-2.5)).tobytes()


def test_tee_reference_test_simple():
    """src/tee.rs `simple`: 10 floats, one work() -> WaitForStream(_, 1), both sides equal the input."""
    samps = np.arange(10, dtype=np.float32)
    vs = B.VectorSource(samps)
    vs.work()
    tee = B.Tee(vs.out)
    r = tee.work()
    assert r.kind == B.WAIT and r.need == 1
    for o in (tee.out1, tee.out2):
        got, tags = o.read_buf()
        assert got.tolist() == samps.tolist()
        assert [t.key for t in tags] == ["VectorSource::start", "VectorSource::repeat", "VectorSource::first"]


def test_iq_balance_reference_test():
    """src/iq_balance.rs `removes_dc_offset_quickly_with_large_alpha`."""
    src = B.Stream(np.complex64)
    fill(src, np.full(8, 1.0 - 2.0j, np.complex64))
    b = B.IqBalance(src, 0.5)
    b.work()
    s, _ = b.out.read_buf()
    assert len(s) == 8 and abs(s[-1].real) < 0.01 and abs(s[-1].imag) < 0.01
    # exact: residual after n steps is (1 - alpha)^n * x
    assert s[-1] == np.complex64((1.0 - 2.0j) * 0.5 ** 8)
    # alpha is clamped to [0, 1] (with_alpha, :63)
    assert O.IqBalance(7.0).alpha == 1.0 and O.IqBalance(-1.0).alpha == 0.0
    x = O.synth_c32(34, 0, 5000) + np.complex64(0.25 + 0.5j)
    e = O.IqBalance(0.01)
    assert O.rel_rms(e.work(x), O.IqBalance(0.01).work(x, f64=True)) < 1e-6
    assert abs(e.mean[0] - (0.25 + 0.5j)) < 0.2


def test_sync_blockmodel_partial_output_space():
    """Sync loop: n = min(input, output space); tags beyond n stay for the next round."""
    src = B.Stream(np.float32)
    x = O.synth_f32(35, 0, 100)
    fill(src, x, [B.tag_u64(5, "k", 1), B.tag_u64(80, "k", 2)])
    m = B.MultiplyConst(src, 2.0, stream_bytes=64 * 4)      # room for 64 samples
    r = m.work()
    assert r.kind == B.WAIT and r.stream is m.out and r.need == 1
    got, tags = m.out.read_buf()
    assert len(got) == 64 and tags == [B.tag_u64(5, "k", 1)]
    m.out.consume(64)
    r = m.work()
    assert r.kind == B.WAIT and r.stream is src
    got, tags = m.out.read_buf()
    assert len(got) == 36 and tags == [B.tag_u64(16, "k", 2)]


def test_no_cpu_fallback_for_the_new_entry_points():
    import rustradio_b200 as R
    try:
        have = R.device_count() > 0
    except R.RrcError:
        have = False
    if have:
        pytest.skip("box has a GPU")
    with pytest.raises(R.RrcError):
        R.Hilbert(65)
    with pytest.raises(R.RrcError):
        R.IqBalance(0.5)
    with pytest.raises(R.RrcError):
        R.multiply_const(np.ones(8, np.float32), 2.0)


# ------------------------------------------------------------------ GPU ---
@pytest.fixture(scope="module")
def R():
    import rustradio_b200 as R
    assert R.device_count() >= 1
    return R


@pytest.mark.gpu
@pytest.mark.parametrize("ntaps,wt", [(65, 0), (3, 0), (127, 1), (1001, 2), (8191, 0)])
def test_hilbert_kernel_vs_oracle(R, ntaps, wt):
    n = 200_003
    x = O.synth_f32(40 + ntaps, 0, n)
    want64 = O.Hilbert(ntaps, wt).work(x, f64=True)
    h = R.Hilbert(ntaps, wt)
    got = h.process(x)
    assert len(got) == n
    assert got.real.tobytes() == want64.real.astype(np.float32).tobytes()      # delayed copy: exact
    assert O.rel_rms(got.imag, want64.imag) <= 1e-5                              # bar: 1e-5 rel-RMS
    assert O.rel_rms(got, O.Hilbert(ntaps, wt).work(x)) <= 1e-5
    # chunking independence incl. pieces shorter than the history and an empty call
    h2 = R.Hilbert(ntaps, wt)
    cuts = [0, 1, 3, 3, 70, 5000, 5001, 20_000, 150_001, n]
    pieces = [h2.process(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.concatenate(pieces).tobytes() == got.tobytes()
    h2.reset()
    assert h2.process(x[:1000]).tobytes() == got[:1000].tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("ntaps", [3, 65, 67, 129])
def test_hilbert_dense_kernel_for_arbitrary_taps(R, ntaps):
    """Taps that are not half-band (caller supplied through rrc_hilbert_create) take hilbert_kernel; the
    reference's own taps take hilbert_half_kernel (both parities of T/2: 65 -> odd taps, 67 -> even)."""
    n = 50_001
    x = O.synth_f32(70 + ntaps, 0, n)
    taps = O.synth_f32(71, 0, ntaps) / np.float32(ntaps)
    got = R.Hilbert(ntaps, taps=taps).process(x)
    assert O.rel_rms(got, O.Hilbert(ntaps, taps=taps).work(x, f64=True)) <= 1e-5
    got = R.Hilbert(ntaps).process(x)
    assert O.rel_rms(got, O.Hilbert(ntaps).work(x, f64=True)) <= 1e-5


@pytest.mark.gpu
def test_hilbert_constructor_errors(R):
    for bad in (0, 1, 64):
        with pytest.raises(R.RrcError):
            R.Hilbert(bad)
    with pytest.raises(R.RrcError):
        R.Hilbert(8193)


@pytest.mark.gpu
@pytest.mark.parametrize("n,off", [(0, 0), (1, 0), (3, 0), (4, 0), (1_000_003, 0), (100_001, 1), (4099, 3)])
def test_maps_bit_exact(R, n, off):
    xc = O.synth_c32(50, 0, n)
    xf = O.synth_f32(51, 0, n)
    v = 0.3 - 1.7j
    assert R.multiply_const(xc, v, offset=off).tobytes() == O.multiply_const(xc, v).tobytes()
    assert R.add_const(xc, v, offset=off).tobytes() == O.add_const(xc, v).tobytes()
    assert R.multiply_const(xf, 0.37, offset=off).tobytes() == O.multiply_const(xf, 0.37).tobytes()
    assert R.add_const(xf, -2.5, offset=off).tobytes() == O.add_const(xf, -2.5).tobytes()
    assert R.complex_to_mag2(xc, offset=off).tobytes() == O.complex_to_mag2(xc).tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("nbytes,off", [(0, 0), (1, 0), (15, 0), (16, 0), (4_000_001, 0), (100_000, 4), (100_004, 8), (77_777, 3)])
def test_tee_kernel(R, nbytes, off):
    x = O.synth_u8(52, 0, nbytes)
    a, b = R.tee(x, offset_bytes=off)
    assert a.tobytes() == x.tobytes() and b.tobytes() == x.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("alpha", [0.5, 0.01, 1e-4, 1.0, 0.0])
@pytest.mark.parametrize("n", [1, 8, 4095, 4096, 4097, 300_001, 5_000_000])
def test_iq_balance_kernel_vs_oracle(R, alpha, n):
    x = (O.synth_c32(53, 0, n) + np.complex64(0.25 - 0.5j)).astype(np.complex64)
    ref = O.IqBalance(alpha)
    want = ref.work(x)
    want64 = O.IqBalance(alpha).work(x, f64=True)
    b = R.IqBalance(alpha)
    got = b.process(x)
    assert O.rel_rms(got, want64) <= 1e-5             # bar: 1e-5 rel-RMS vs the f64 recurrence
    assert O.rel_rms(got, want) <= 1e-5               # and vs the reference's sequential f32 loop
    assert abs(b.mean - complex(ref.mean[0])) <= 1e-5 * max(1.0, abs(ref.mean[0]))
    # carried mean: a second call continues the same stream
    x2 = O.synth_c32(54, 0, 10_000)
    assert O.rel_rms(b.process(x2), ref.work(x2)) <= 1e-5
    # slicing independence (tolerance: the association changes with the tiling)
    if n > 4097:
        b2 = R.IqBalance(alpha)
        cuts = [0, 1, 4096, 10_000, n]
        pieces = [b2.process(x[a:c]) for a, c in zip(cuts[:-1], cuts[1:])]
        assert O.rel_rms(np.concatenate(pieces), want64) <= 1e-5


@pytest.mark.gpu
def test_iq_balance_reference_test_gpu(R):
    got = R.IqBalance(0.5).process(np.full(8, 1.0 - 2.0j, np.complex64))
    assert abs(got[-1].real) < 0.01 and abs(got[-1].imag) < 0.01


def _run_chain(K, data, build, residency):
    src, s0 = K.VectorSource(data, residency=residency)
    blk, out = build(s0)
    K.graph_run([src, blk])
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("res", ["DEVICE", "HOST"])
def test_blocks_vs_blockmodel(R, res):
    """Every new rr:: block driven by Graph::run against the block model: samples, counts, tags."""
    from rustradio_b200 import blocks as K
    r = getattr(K, res)
    n = 200_000
    xf = O.synth_f32(60, 0, n)
    xc = O.synth_c32(61, 0, n)
    vtags = [K.Tag(0, "VectorSource::start", ("Bool", True)), K.Tag(0, "VectorSource::repeat", ("U64", 0)),
             K.Tag(0, "VectorSource::first", ("Bool", True))]

    out = _run_chain(K, xf, lambda s: K.Hilbert(s, 65, residency=r), r)
    got, tags = out.read_buf()
    assert len(got) == n and tags == vtags
    assert O.rel_rms(got, O.Hilbert(65).work(xf, f64=True)) <= 1e-5

    for data, val in ((xf, 0.37), (xc, 0.3 - 1.7j)):
        out = _run_chain(K, data, lambda s: K.MultiplyConst(s, val, residency=r), r)
        got, tags = out.read_buf()
        assert got.tobytes() == O.multiply_const(data, val).tobytes() and tags == vtags
        out = _run_chain(K, data, lambda s: K.AddConst(s, val, residency=r), r)
        got, tags = out.read_buf()
        assert got.tobytes() == O.add_const(data, val).tobytes() and tags == vtags

    out = _run_chain(K, xc, lambda s: K.ComplexToMag2(s, residency=r), r)
    got, tags = out.read_buf()
    assert got.dtype == np.float32 and got.tobytes() == O.complex_to_mag2(xc).tobytes() and tags == vtags

    out = _run_chain(K, xc, lambda s: K.IqBalance(s, 0.01, residency=r), r)
    got, tags = out.read_buf()
    assert len(got) == n and tags == vtags
    assert O.rel_rms(got, O.IqBalance(0.01).work(xc)) <= 1e-5

    src, s0 = K.VectorSource(xc, residency=r)
    tee, o1, o2 = K.Tee(s0, residency=r)
    assert tee.name == "Tee"
    K.graph_run([src, tee])
    for o in (o1, o2):
        got, tags = o.read_buf()
        assert got.tobytes() == xc.tobytes() and tags == vtags


@pytest.mark.gpu
def test_tee_reference_test_simple_gpu(R):
    """src/tee.rs `simple` against the rr::Tee block."""
    from rustradio_b200 import blocks as K
    samps = np.arange(10, dtype=np.float32)
    src, s0 = K.VectorSource(samps)
    src.work()
    tee, o1, o2 = K.Tee(s0)
    ret = tee.work()
    assert ret.kind == K.WAIT and ret.need == 1
    for o in (o1, o2):
        got, _ = o.read_buf()
        assert got.tolist() == samps.tolist()


@pytest.mark.gpu
def test_sync_block_partial_output_space(R):
    """n = min(input, output space); a tag beyond n is delivered by the next round at its re-based position."""
    from rustradio_b200 import blocks as K
    w, r = K.new_stream(np.float32, residency=K.HOST)
    x = O.synth_f32(35, 0, 100)
    w.write(x, [K.Tag(5, "k", ("U64", 1)), K.Tag(80, "k", ("U64", 2))])
    m, out = K.MultiplyConst(r, 2.0, size_bytes=4096, residency=K.HOST)      # one page: room for 1024 samples
    pre = 1024 - 64                                                          # leave room for 64 samples
    # occupy the output ring through the block itself: feed `pre` samples first
    w.write(np.zeros(0, np.float32))
    ret = m.work()
    got, tags = out.read_buf()
    assert ret.kind == K.WAIT and len(got) == 100 and tags == [K.Tag(5, "k", ("U64", 1)), K.Tag(80, "k", ("U64", 2))]
    assert got.tobytes() == (x * np.float32(2.0)).tobytes()
    # now fill the output so that only 64 samples of space remain, and offer 100 more with tags
    w.write(np.zeros(pre - 100, np.float32))
    m.work()
    assert len(out.read_buf()[0]) == pre
    w.write(x, [K.Tag(5, "k", ("U64", 1)), K.Tag(80, "k", ("U64", 2))])
    ret = m.work()
    assert ret.kind == K.WAIT and ret.stream_id == out.id and ret.need == 1
    got, tags = out.read_buf()
    assert len(got) == 1024 and [t for t in tags if t.pos >= pre] == [K.Tag(pre + 5, "k", ("U64", 1))]
    assert got[pre:].tobytes() == (x[:64] * np.float32(2.0)).tobytes()
    out.consume(1024)
    ret = m.work()
    assert ret.kind == K.WAIT and ret.stream_id != out.id
    got, tags = out.read_buf()
    assert len(got) == 36 and tags == [K.Tag(16, "k", ("U64", 2))]


@pytest.mark.gpu
def test_graph_hilbert_chain(R):
    """examples/ax25-1200-rx.rs shape: f32 audio -> Hilbert(65, Hamming) -> QuadratureDemod."""
    from rustradio_b200 import blocks as K
    n = 300_000
    t = np.arange(n)
    x = (np.cos(2 * np.pi * 0.05 * t) + 0.01 * O.synth_f32(62, 0, n)).astype(np.float32)
    src, s0 = K.VectorSource(x)
    hil, s1 = K.Hilbert(s0, 65)
    dem, s2 = K.QuadratureDemod(s1, 1.0)
    K.graph_run([src, hil, dem])
    got, _ = s2.read_buf()
    assert len(got) == n - 1
    y = O.Hilbert(65).work(x, f64=True)
    assert O.max_angle_err(got[200:], np.angle(y[201:] * np.conj(y[200:-1]))) <= 1e-4


# ------------------------------------------------ neighbours as store epilogues ---
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(64, 1), (255, 10), (130, 5), (33, 4), (5, 1), (255, 3)])
def test_fir_store_epilogues_equal_the_unfused_chain(R, shape):
    """FirFilter -> MultiplyConst / AddConst / ComplexToMag2 fused into the filter's store equals
    the filter followed by the stand-alone block BIT FOR BIT (same FP32 kernel, same separately rounded ops);
    covers the uniform-tap kernel (255/10, 130/5), the packed kernel, the FP32 poly kernel and complex taps."""
    ntaps, deci = shape
    x = O.synth_c32(71, 0, 120_000)
    for cplx_taps in (False, True):
        taps = O.low_pass_n(1.0, 0.08, ntaps).astype(np.complex64) * ((1 - 0.4j) if cplx_taps else 1)
        base = R.Fir(taps, deci=deci, flags=R.RRC_FIR_NO_TENSOR).filter(x)          # the FP32 kernels, unfused
        for kind, val, ref in ((R.EPI_MULTIPLY_CONST, 0.3 - 1.7j, lambda y: R.multiply_const(y, 0.3 - 1.7j)),
                               (R.EPI_ADD_CONST, -0.25 + 2j, lambda y: R.add_const(y, -0.25 + 2j)),
                               (R.EPI_MAG2, 0, lambda y: R.complex_to_mag2(y))):
            f = R.Fir(taps, deci=deci)
            f.set_epilogue(kind, val)
            assert not f.uses_tensor_cores
            n_out = f.out_count(len(x))
            need = (n_out - 1) * deci + ntaps
            din = R.DeviceBuffer.from_numpy(x[:need])
            mag = kind == R.EPI_MAG2
            dout = R.DeviceBuffer(n_out * (4 if mag else 8))
            f.run(din, need, dout, n_out)
            got = dout.download(np.float32 if mag else np.complex64, n_out)
            want = ref(base)
            assert got.tobytes() == want.tobytes(), (shape, cplx_taps, kind)
    f = R.Fir(taps, deci=deci)
    f.set_epilogue(R.EPI_MULTIPLY_CONST, 2.0)
    with pytest.raises(R.RrcError):                              # demod's gain IS the fused MultiplyConst
        f.demod_run_batch(din, len(x), need, 1.0, dout, n_out - 1, n_out, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("ntaps", [257, 4097])
def test_fftfilter_store_epilogues_equal_the_unfused_chain(R, ntaps):
    """FftFilter -> MultiplyConst / ComplexToMag2 fused into phase A' / the fold kernel's combine store equals the two
    blocks back to back bit for bit, for the plain run, the store-predicate decimation and the decimate-by-8 kernel."""
    x = O.synth_c32(72, 0, 150_000)
    taps = O.low_pass_n(1.0, 0.05, ntaps).astype(np.complex64)
    dx = R.DeviceBuffer.from_numpy(x)
    n = len(x)
    y = R.FftFilt(taps).filter(x)
    for kind, val, ref in ((R.EPI_MULTIPLY_CONST, 0.5 + 0.25j, lambda v: R.multiply_const(v, 0.5 + 0.25j)),
                           (R.EPI_MAG2, 0, lambda v: R.complex_to_mag2(v))):
        mag = kind == R.EPI_MAG2
        dt, esz = (np.float32, 4) if mag else (np.complex64, 8)
        f = R.FftFilt(taps)
        f.set_epilogue(kind, val)
        dout = R.DeviceBuffer(n * esz)
        f.run(dx, n, dout)
        assert dout.download(dt, n).tobytes() == ref(y).tobytes(), (ntaps, kind, "run")
        for deci, skip in ((8, 0), (8, 5), (3, 1)):
            plain = R.FftFilt(taps)
            cnt = (n - skip + deci - 1) // deci
            d0 = R.DeviceBuffer(cnt * 8)
            assert plain.decim_run(dx, n, deci, skip, d0) == cnt
            f = R.FftFilt(taps)
            f.set_epilogue(kind, val)
            d1 = R.DeviceBuffer(cnt * esz)
            assert f.decim_run(dx, n, deci, skip, d1) == cnt
            assert d1.download(dt, cnt).tobytes() == ref(d0.download(np.complex64, cnt)).tobytes(), (ntaps, kind, deci, skip)
    # end to end: the MAG2 form halves the D2H bytes
    f = R.FftFilt(taps)
    f.set_epilogue(R.EPI_MAG2)
    out = np.empty((n // f.nsamples) * f.nsamples, np.float32)
    got = f.run_host(x, out)
    assert got.tobytes() == R.complex_to_mag2(R.FftFilt(taps).run_host(x)).tobytes()
