"""Capture ingest (second half): FileSource / SigMFSource into host, pinned-host and
device rings.  The reference's own #[test]s (src/file_source.rs:160-287, src/sigmf.rs:616-631) are
restated one by one; the HOST-ring variants need no GPU (the file I/O is host work) and run in the CPU
suite, the DEVICE / HOST_PINNED variants and the file -> GPU filter chains are marked gpu."""
import io
import json
import tarfile

import numpy as np
import pytest

from oracle import oracle as O


@pytest.fixture(scope="module")
def K():
    from rustradio_b200 import blocks as K
    return K


def _gpu_available():
    try:
        import rustradio_b200 as R
        return R.device_count() >= 1
    except Exception:
        return False


RES = [pytest.param("HOST"), pytest.param("DEVICE", marks=pytest.mark.gpu), pytest.param("HOST_PINNED", marks=pytest.mark.gpu)]
F32_BYTES = bytes([0, 0, 128, 63, 0, 0, 64, 64, 195, 245, 72, 64, 195, 245, 72, 192])


@pytest.mark.parametrize("res", RES)
def test_source_dst_full(K, res):
    """src/file_source.rs:160-168: /dev/zero fills the stream (Again), then WaitForStream(dst, 1)."""
    src, out = K.FileSource("/dev/zero", np.float32, residency=getattr(K, res))
    assert src.work().kind == K.AGAIN
    ret = src.work()
    assert (ret.kind, ret.stream_id, ret.need) == (K.WAIT, out.id, 1)
    assert len(out) == out.capacity


@pytest.mark.parametrize("res", RES)
def test_source_f32_and_partial_tail_and_twice(K, res, tmp_path):
    """src/file_source.rs:170-252."""
    r = getattr(K, res)
    fn = tmp_path / "delme.bin"
    fn.write_bytes(F32_BYTES)
    src, out = K.FileSource(fn, np.float32, residency=r)
    assert src.work().kind == K.AGAIN
    assert src.work().kind == K.EOF
    got, tags = out.read_buf()
    assert got.tobytes() == np.array([1.0, 3.0, 3.14, -3.14], np.float32).tobytes() and tags == []
    fn.write_bytes(F32_BYTES[:-1])                               # source_f32_partial_tail
    src, out = K.FileSource(fn, np.float32, residency=r)
    assert src.work().kind == K.AGAIN
    assert src.work().kind == K.EOF
    assert out.read_buf()[0].tobytes() == np.array([1.0, 3.0, 3.14], np.float32).tobytes()
    fn.write_bytes(F32_BYTES)                                    # source_f32_twice
    src, out = K.FileSource(fn, np.float32, repeat=2, residency=r)
    assert [src.work().kind for _ in range(4)] == [K.AGAIN, K.AGAIN, K.AGAIN, K.EOF]
    assert out.read_buf()[0].tobytes() == np.array([1.0, 3.0, 3.14, -3.14] * 2, np.float32).tobytes()


@pytest.mark.parametrize("res", RES)
def test_source_repeat_discards_partial_tail_and_c32(K, res, tmp_path):
    """src/file_source.rs:254-287."""
    r = getattr(K, res)
    fn = tmp_path / "delme.bin"
    fn.write_bytes(bytes([1, 0, 0, 0, 0xff]))
    src, out = K.FileSource(fn, np.uint32, repeat=2, residency=r)
    for _ in range(100):
        if src.work().kind == K.EOF:
            break
    else:
        raise AssertionError("no EOF")
    assert list(out.read_buf()[0]) == [1, 1]
    fn.write_bytes(bytes([0, 0, 0, 0, 0, 0, 0, 0, 195, 245, 72, 64, 205, 204, 44, 192]))
    src, out = K.FileSource(fn, np.complex64, residency=r)
    src.work()
    assert out.read_buf()[0].tobytes() == np.array([0, 3.14 - 2.7j], np.complex64).tobytes()


def test_source_missing_file_is_an_error(K):
    """File::open error -> Err(Error::file_io) (src/file_source.rs:64-66)."""
    from rustradio_b200 import RrcError
    with pytest.raises(RrcError):
        K.FileSource("/nonexistent/definitely/not/here.cf32", np.complex64, residency=K.HOST)


def _write_recording(base, data: bytes, datatype="cf32_le", rate=None):
    g = {"core:datatype": datatype, "core:version": "1.1.0"}
    if rate is not None:
        g["core:sample_rate"] = rate
    (base.parent / (base.name + "-meta")).write_text(json.dumps({"global": g, "captures": [], "annotations": []}))
    (base.parent / (base.name + "-data")).write_bytes(data)


@pytest.mark.parametrize("res", RES)
def test_sigmf_partial_tail_makes_progress_to_eof(K, res, tmp_path):
    """src/sigmf.rs:617-629."""
    base = tmp_path / "partial"
    _write_recording(base, bytes([0xff]), "rf32_le")
    src, out, rate = K.SigMFSource(base, np.float32, residency=getattr(K, res))
    assert src.work().kind == K.EOF
    assert len(out) == 0 and rate is None


@pytest.mark.parametrize("res", RES)
def test_sigmf_recording_archive_types_and_rates(K, res, tmp_path):
    """Recording files and the tar Archive form give the same samples; datatype / sample-rate checks of
    new2 (src/sigmf.rs:389-412); work() returns WaitForStream(dst, 1) after each produce (:609)."""
    from rustradio_b200 import RrcError
    r = getattr(K, res)
    x = O.synth_c32(91, 0, 70_000)
    base = tmp_path / "cap.sigmf"
    _write_recording(base, x.tobytes() + b"\x01\x02\x03", "cf32_le", rate=2_400_000.0)    # 3 stray tail bytes
    src, out, rate = K.SigMFSource(base, np.complex64, sample_rate=2_400_000.0, repeat=2, size_bytes=4096 * 300, residency=r)
    assert rate == 2_400_000.0
    kinds = []
    for _ in range(20):
        ret = src.work()
        kinds.append(ret.kind)
        if ret.kind == K.EOF:
            break
    assert kinds[-1] == K.EOF and set(kinds[:-1]) == {K.WAIT}
    got = out.read_buf()[0]
    assert got.tobytes() == np.concatenate([x, x]).tobytes()
    with pytest.raises(RrcError):                                 # sample-rate mismatch
        K.SigMFSource(base, np.complex64, sample_rate=1_000_000.0, residency=r)
    with pytest.raises(RrcError):                                 # type mismatch: file is cf32_le
        K.SigMFSource(base, np.float32, residency=r)
    src, out, _ = K.SigMFSource(base, np.float32, ignore_type_error=True, residency=r)
    src.work()
    assert out.read_buf(10)[0].tobytes() == x[:5].tobytes()
    with pytest.raises(RrcError):                                 # neither archive nor recording files
        K.SigMFSource(tmp_path / "nothing.sigmf", np.complex64, residency=r)
    # Archive: a tar holding rec.sigmf-meta and rec.sigmf-data
    arc = tmp_path / "arc.sigmf"
    with tarfile.open(arc, "w", format=tarfile.USTAR_FORMAT) as tf:
        meta = json.dumps({"global": {"core:datatype": "cf32_le", "core:sample_rate": 48000}, "captures": []}).encode()
        for name, payload in (("rec/rec.sigmf-meta", meta), ("rec/rec.sigmf-data", x.tobytes())):
            ti = tarfile.TarInfo(name)
            ti.size = len(payload)
            tf.addfile(ti, io.BytesIO(payload))
    src, out, rate = K.SigMFSource(arc, np.complex64, size_bytes=4096 * 150, residency=r)
    assert rate == 48000.0
    while src.work().kind != K.EOF:
        pass
    assert out.read_buf()[0].tobytes() == x.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("res", ["DEVICE", "HOST_PINNED", "HOST"])
def test_capture_file_through_the_gpu_filters(K, res, tmp_path):
    """A cf32 capture on disk -> FileSource -> FftFilter -> RationalResampler(1, 8) -> host, through device
    rings (config 5's chain from its on-disk entry) equals the oracle chain."""
    r = getattr(K, res)
    n, T = 300_000, 1025
    x = O.synth_c32(92, 0, n)
    fn = tmp_path / "capture.cf32"
    fn.write_bytes(x.tobytes())
    taps = O.low_pass_n(1.0, 0.05, T).astype(np.complex64)
    src, s1 = K.FileSource(fn, np.complex64, size_bytes=1 << 20, residency=r)
    flt, s2 = K.FftFilter(s1, taps, size_bytes=1 << 20, residency=K.DEVICE)
    rs, s3 = K.RationalResampler(s2, 1, 8, size_bytes=1 << 20, residency=r)
    outs = []
    for _ in range(10_000):
        kinds = [b.work().kind for b in (src, flt, rs)]
        if len(s3):
            outs.append(s3.read_buf()[0])
            s3.consume(len(outs[-1]))
        elif kinds[0] == K.EOF and kinds[1] == K.WAIT and kinds[2] == K.WAIT:
            break
    got = np.concatenate(outs)
    n_filt = O.fftfilt_out_count(n, T)
    want = O.conv_full_f64_fft(x, taps, n_filt)[::8]
    assert len(got) == len(want) and O.rel_rms(got, want) <= 1e-5


@pytest.mark.gpu
def test_pinned_host_ring_roundtrip_with_tags(K):
    """HOST_PINNED rings behave like HOST rings (same counts / tags), with the block's copies running as DMA
    straight from / to the ring windows, including windows that wrap through the doubled mapping."""
    x = O.synth_c32(93, 0, 200_000)
    taps = O.low_pass_n(1.0, 0.1, 33).astype(np.complex64)
    want = O.fir(x, taps, 3, f64=True)
    for res in (K.HOST_PINNED, K.HOST):
        w, r = K.new_stream(np.complex64, size_bytes=4096 * 16, residency=res)
        blk, out = K.FirFilter(r, taps, 3, size_bytes=4096 * 16, residency=res)
        got, tags, fed, produced = [], [], 0, 0
        while True:
            m = min(w.free(), len(x) - fed)
            if m:
                t = [K.Tag(p - fed, "t", ("U64", p)) for p in range(-(-fed // 7919) * 7919, fed + m, 7919)]
                assert w.write(x[fed:fed + m], t) == m
                fed += m
            progressed = False
            while blk.work().kind == K.AGAIN:
                progressed = True
            if len(out):
                y, tg = out.read_buf()
                tags += [(produced + q.pos, q.val[1]) for q in tg]
                got.append(y)
                produced += len(y)
                out.consume(len(y))
                progressed = True
            if fed == len(x) and not progressed:
                break
        y = np.concatenate(got)
        assert len(y) == len(want) and O.rel_rms(y, want) <= 1e-5
        assert tags == [(p // 3, p) for p in range(0, len(x), 7919) if p < len(want) * 3]
