"""Block-level parity on the GPU: the C++ mirror of rustradio's Block contract
(rrb_* ABI -> csrc/blocks.cu -> CUDA kernels) driven exactly like the
reference's own #[test]s drive its blocks: VectorSource -> work() by hand ->
assert BlockRet variants -> read_buf() and compare samples AND tags.
Each test names the reference test it restates; expected BlockRet sequences,
counts and tags are additionally cross-checked against the restated CPU model
(oracle/blockmodel.py)."""
import numpy as np
import pytest

from oracle import blockmodel as B
from oracle import oracle as O

pytestmark = pytest.mark.gpu

INPUT6 = np.array([1, 2, 3 + .2j, 4.1, 5, 6 + .2j], np.complex64)


@pytest.fixture(scope="module")
def K():
    import rustradio_b200 as R
    from rustradio_b200 import blocks as K
    assert R.device_count() >= 1
    return K


def almost(a, b, tol=1e-3):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a, b)
    assert np.all(np.abs(a - b) <= tol), (a, b)


def vs_tags_2rep(K, p):
    return [K.Tag(0, "VectorSource::start", ("Bool", True)), K.Tag(0, "VectorSource::repeat", ("U64", 0)),
            K.Tag(0, "VectorSource::first", ("Bool", True)), K.Tag(p, "VectorSource::start", ("Bool", True)),
            K.Tag(p, "VectorSource::repeat", ("U64", 1))]


@pytest.mark.parametrize("res", ["DEVICE", "HOST"])
def test_fir_test_identity(K, res):
    """src/fir.rs:691-741: counts, tags at 0 and 6/deci, Again then WaitForStream."""
    r = getattr(K, res)
    for deci in range(1, 3 * len(INPUT6) + 1):
        src, so = K.VectorSource(INPUT6, repeat=2, residency=r)
        assert src.work().kind == K.AGAIN
        assert src.work().kind == K.EOF
        b, os_ = K.FirFilter(so, np.array([1], np.complex64), deci, residency=r)
        if deci <= 2 * len(INPUT6):
            assert b.work().kind == K.AGAIN
        ret = b.work()
        assert ret.kind == K.WAIT
        res_, tags = os_.read_buf()
        mx = 2 * len(INPUT6) // deci
        if len(res_):
            assert tags == vs_tags_2rep(K, 6 // deci)
        almost(res_, np.concatenate([INPUT6, INPUT6])[::deci][:mx])


def test_fir_moving_avg_and_invert(K):
    """src/fir.rs:843-919"""
    want = np.array([1.5, 2.5 + .1j, 3.55 + .1j, 4.55, 5.5 + .1j], np.complex64)
    for deci in range(1, len(INPUT6) + 2):
        src, so = K.VectorSource(INPUT6)
        src.work()
        b, os_ = K.FirFilter(so, np.array([.5, .5], np.complex64), deci)
        if deci < len(INPUT6):
            assert b.work().kind == K.AGAIN
        assert b.work().kind == K.WAIT
        almost(os_.read_buf()[0], want[::deci][:(len(INPUT6) - 1) // deci])
        src, so = K.VectorSource(INPUT6)
        src.work()
        b, os_ = K.FirFilter(so, np.array([-1], np.complex64), deci)
        if deci <= len(INPUT6):
            assert b.work().kind == K.AGAIN
        assert b.work().kind == K.WAIT
        almost(os_.read_buf()[0], -INPUT6[::deci][:len(INPUT6) // deci])


def test_fir_wait_identifies_stream_and_need(K):
    """src/fir.rs:496-515: WaitForStream(src, ntaps+deci-1) / WaitForStream(dst, 1)."""
    w, r = K.new_stream(np.complex64, residency=K.DEVICE)
    b, out = K.FirFilter(r, np.ones(5, np.complex64), 3)
    ret = b.work()
    assert (ret.kind, ret.stream_id, ret.need) == (K.WAIT, w.id, 5 + 3 - 1)
    assert b.name == "FirFilter<Complex>"
    # fill the output completely -> WaitForStream(dst, 1)
    cap = out.capacity
    w.write(np.ones(4096, np.complex64))
    while b.work().kind == K.AGAIN:
        pass
    n_out_total = 0
    while True:
        free = w.free()
        w.write(np.ones(min(free, 100_000), np.complex64))
        ret = b.work()
        if ret.kind == K.WAIT and ret.stream_id == out.id:
            break
    assert ret.need == 1 and len(out) == cap


def test_fir_translate_matches_mixed_input(K):
    """src/fir.rs:744-789"""
    inp = np.array([complex(i, i * 0.25) for i in range(32)], np.complex64)
    taps = np.array([.5 - .1j, 1 + .2j, -.25 + .05j, .125 - .3j], np.complex64)
    phase_step = -2.0 * np.pi * 2.0 / 8.0
    rot = np.complex64(complex(np.float32(np.cos(phase_step)), np.float32(np.sin(phase_step))))
    phase = np.complex64(1)
    mixed = np.empty_like(inp)
    for i, s in enumerate(inp):
        mixed[i] = np.complex64(s * phase)
        phase = np.complex64(phase * rot)
    sa, sao = K.VectorSource(inp)
    assert sa.work().kind == K.EOF
    tr, tro = K.FirFilter(sao, taps, 3, translate=(8.0, 2.0))
    assert tr.work().kind == K.AGAIN and tr.work().kind == K.WAIT
    sb, sbo = K.VectorSource(mixed)
    assert sb.work().kind == K.EOF
    man, mano = K.FirFilter(sbo, taps, 3)
    assert man.work().kind == K.AGAIN and man.work().kind == K.WAIT
    a, _ = tro.read_buf()
    b, _ = mano.read_buf()
    assert len(a) == len(b) == 9
    almost(a, b)


def test_fir_chunked_stream_equals_whole_stream_with_tags(K):
    """Appendix A: output and tag positions are independent of how work() slices the stream."""
    x = O.synth_c32(3, 0, 60_000)
    taps = O.low_pass_complex(48000.0, 3000.0, 2000.0)
    deci = 3
    whole = O.fir(x, taps, deci, f64=True)
    w, r = K.new_stream(np.complex64, residency=K.DEVICE)
    f, out = K.FirFilter(r, taps, deci)
    got, got_tags, pos, opos = [], [], 0, 0
    rng = np.random.default_rng(0)
    while pos < len(x):
        n = min(int(rng.integers(1, 9000)), len(x) - pos)
        assert w.write(x[pos:pos + n], [K.Tag(0, "chunk", ("U64", pos))]) == n
        pos += n
        while f.work().kind == K.AGAIN:
            pass
        o, tags = out.read_buf()
        got.append(o)
        got_tags += [(opos + t.pos, t.val[1]) for t in tags]
        opos += len(o)
        out.consume(len(o))
    got = np.concatenate(got)
    assert len(got) == len(whole)
    assert O.rel_rms(got, whole) <= 1e-5
    # a tag at absolute input p < M*D lands on output floor(p/D)
    assert all(op == p // deci for op, p in got_tags)
    assert len(got_tags) >= 1 and all(p < len(whole) * deci for _, p in got_tags)


@pytest.mark.parametrize("res", ["DEVICE", "HOST"])
def test_fftfilter_tag_propagation(K, res):
    """src/fft_filter.rs:551-574"""
    r = getattr(K, res)
    src, so = K.VectorSource(np.zeros(1024, np.complex64), repeat=2, residency=r)
    f, out = K.FftFilter(so, np.zeros(1, np.complex64), residency=r)
    src.work()
    src.work()
    f.work()
    o, tags = out.read_buf()
    assert tags == vs_tags_2rep(K, 1024)
    assert len(o) == 2048


def test_fftfilter_filter_a_signal(K):
    """src/fft_filter.rs:502-549 + BlockRet/need from the restated model."""
    taps = O.low_pass_complex(8000.0, 1000.0, 100.0)
    sig, _ = O.signal_source_complex(8000.0, 3000.0, 1.0, 8000)
    src, so = K.VectorSource(sig)
    src.work()
    f, out = K.FftFilter(so, taps)
    ret = f.work()
    assert (ret.kind, ret.need) == (K.WAIT, 319 - (8000 - 25 * 319))
    o, tags = out.read_buf()
    assert len(o) == 7975
    assert np.max(np.abs(o[len(taps):])) < 2e-4
    assert all(t.pos == 0 for t in tags)


def test_fftfilter_partial_blocks_and_tags_match_model(K):
    """Random chunking: consumed counts, BlockRet, emitted samples and tag positions equal the restated work()."""
    taps = O.low_pass_n(1.0, 0.1, 300).astype(np.complex64)     # fft 1024, block 724
    x = O.synth_c32(9, 0, 30_000)
    w, r = K.new_stream(np.complex64, residency=K.DEVICE)
    f, out = K.FftFilter(r, taps)
    ms = B.Stream(np.complex64)
    mf = B.FftFilter(ms, taps)
    rng = np.random.default_rng(1)
    pos, got = 0, []
    while pos < len(x):
        n = min(int(rng.integers(1, 2500)), len(x) - pos)
        tag_g = [K.Tag(int(rng.integers(0, n)), "t", ("U64", pos))]
        w.write(x[pos:pos + n], tag_g)
        mw = ms.write_buf()
        mw[:n] = x[pos:pos + n]
        ms.produce(n, [B.Tag(tag_g[0].pos, "t", ("U64", pos))])
        pos += n
        rg, rm = f.work(), mf.work()
        assert rg.kind == K.WAIT and rm.kind == B.WAIT
        assert rg.need == rm.need and (rg.stream_id == w.id) == (rm.stream is ms)
        og, tg = out.read_buf()
        om, tm = mf.out.read_buf()
        assert len(og) == len(om)
        assert [(t.pos, t.val) for t in tg] == [(t.pos, t.val) for t in tm]
        got.append(og)
        out.consume(len(og))
        mf.out.consume(len(om))
    got = np.concatenate(got)
    assert len(got) == (len(x) // 724) * 724
    assert O.rel_rms(got, O.conv_full_f64_fft(x, taps, len(got))) <= 1e-5


def test_fftfilter_float(K):
    """src/fft_filter.rs:365-491"""
    taps = O.low_pass(8000.0, 1000.0, 100.0)
    x = O.synth_f32(7, 0, 4000)
    src, so = K.VectorSource(x)
    src.work()
    f, out = K.FftFilterFloat(so, taps)
    ret = f.work()
    assert ret.kind == K.WAIT
    o, _ = out.read_buf()
    assert len(o) == O.fftfilt_out_count(4000, len(taps))
    truth = O.conv_full_f64(x.astype(np.complex64), taps.astype(np.complex64), len(o)).real
    assert O.rel_rms(o, truth) <= 1e-5


def test_resampler_reference_tests(K):
    """src/rational_resampler.rs:224-276,363-373"""
    for deci in range(1, len(INPUT6) + 2):
        src, so = K.VectorSource(INPUT6)
        assert src.work().kind == K.EOF
        b, os_ = K.RationalResampler(so, 1, deci)
        assert b.work().kind == K.WAIT
        res, tags = os_.read_buf()
        assert np.array_equal(res, INPUT6[::deci]) and tags == []
    src, so = K.VectorSource(np.arange(50, dtype=np.uint32))
    src.work()
    b, os_ = K.RationalResampler(so, 25, 64)
    b.work()
    assert list(os_.read_buf()[0]) == [0, 2, 5, 7, 10, 12, 15, 17, 20, 23, 25, 28, 30, 33, 35, 38, 40, 43, 46, 48]
    for n, i, d, final in [(10, 1, 1, 10), (10, 1, 2, 5), (10, 2, 1, 20), (100, 2, 3, 67), (100, 3, 2, 150),
                           (100, 300, 200, 150), (100, 200000, 1024000, 20)]:
        src, so = K.VectorSource(np.arange(n, dtype=np.float32).astype(np.complex64))
        src.work()
        b, os_ = K.RationalResampler(so, i, d)
        b.work()
        assert len(os_.read_buf()[0]) == final
    w, r = K.new_stream(np.float32)
    with pytest.raises(Exception):
        K.RationalResampler(r, 0, 1)


def test_resampler_interpolation_survives_full_output_buffer(K):
    """src/rational_resampler.rs:278-299 (pending sample carried across a full output, custom eof())."""
    cap = K.DEFAULT_STREAM_SIZE // 4
    # device rings round the capacity up to the VMM granularity; use a host output ring for the exact reference size
    boundary = cap // 3
    src, so = K.VectorSource(np.arange(boundary + 1, dtype=np.uint32), residency=K.HOST)
    assert src.work().kind == K.EOF
    src.drop()
    b, os_ = K.RationalResampler(so, 3, 1, residency=K.HOST)
    assert os_.capacity == cap and cap % 3 == 1
    ret = b.work()
    assert (ret.kind, ret.need, ret.stream_id) == (K.WAIT, 1, os_.id)
    first, _ = os_.read_buf()
    assert len(first) == cap and first[cap - 1] == boundary
    os_.consume(cap)
    assert not b.eof()
    ret = b.work()
    assert (ret.kind, ret.need) == (K.WAIT, 1)
    assert list(os_.read_buf()[0]) == [boundary, boundary]
    assert b.eof()


def test_quad_demod_reference_tests(K):
    """src/quadrature_demod.rs:173-264"""
    src, so = K.VectorSource(np.zeros(4, np.complex64))
    src.work()
    b, out = K.QuadratureDemod(so, 1.0)
    ret = b.work()
    assert (ret.kind, ret.need) == (K.WAIT, 2)
    assert list(out.read_buf()[0]) == [0.0, 0.0, 0.0]
    for sign in (-1.0, 1.0):
        src, so = K.VectorSource(np.array([1, 0.707 + sign * 0.707j, sign * 1j, -1], np.complex64))
        src.work()
        b, out = K.QuadratureDemod(so, 1.0)
        b.work()
        almost(out.read_buf()[0], sign * np.array([np.pi / 4, np.pi / 4, np.pi / 2], np.float32))
    # fill_out: 512000 in -> 511999 out, one sample carried (host rings have the reference's exact capacity)
    w, r = K.new_stream(np.complex64, residency=K.HOST)
    b, out = K.QuadratureDemod(r, 1.0, residency=K.HOST)
    cur = 0.0

    def fill():
        nonlocal cur
        n = w.free()
        sig, cur = O.signal_source_complex(1200.0, 100.0, 1.0, n, cur)
        assert w.write(sig) == n
        return n
    assert fill() == 512_000
    b.work()
    assert len(out) == 511_999
    fill()
    b.work()
    assert len(out) == 2 * 511_999
    fill()
    b.work()
    assert len(out) == 2 * 512_000
    o, _ = out.read_buf()
    assert abs(float(np.median(o)) - 2 * np.pi * 100 / 1200) < 1e-4


def test_rtl_fm_chain_through_graph(K):
    """examples/rtl_fm.rs:381-419 shape on device-resident rings: FftFilter -> RationalResampler ->
    QuadratureDemod -> FftFilterFloat, run by the Graph::run restatement; equals the restated CPU chain."""
    fs = 1_024_000.0
    n = 400_000
    t = np.arange(n) / fs
    msg = np.sin(2 * np.pi * 1000 * t)
    x = (np.exp(2j * np.pi * np.cumsum(0.05 * msg)) + 0.01 * O.synth_c32(5, 0, n)).astype(np.complex64)
    taps = O.low_pass_complex(fs, 100_000.0, 20_000.0)
    ataps = O.low_pass(200_000.0, 20_000.0, 10_000.0)
    big = 16 << 20
    src, s0 = K.VectorSource(x, size_bytes=big)
    f1, s1 = K.FftFilter(s0, taps, size_bytes=big)
    rs, s2 = K.RationalResampler(s1, 200_000, 1_024_000, size_bytes=big)
    qd, s3 = K.QuadratureDemod(s2, 1.0, size_bytes=big)
    f2, s4 = K.FftFilterFloat(s3, ataps, size_bytes=big)
    K.graph_run([src, f1, rs, qd, f2])
    got, _ = s4.read_buf()
    # CPU restatement of the same chain (closed forms, Appendix A)
    y1 = O.fftfilt(x, taps)
    y2 = O.resample(y1, 200_000, 1_024_000)
    y3 = O.quad_demod(y2, 1.0)
    y4 = O.fftfilt(y3.astype(np.complex64), ataps.astype(np.complex64)).real
    assert len(got) == len(y4) > 10_000
    assert O.rel_rms(got, y4) < 1e-3      # demod of a noisy carrier amplifies f32 differences; counts are exact
