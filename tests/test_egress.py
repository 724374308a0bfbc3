"""The output side of the path (SURVEY section 2 "sample-format codecs" / 8f "formats either side"): RtlSdrEncode
(Complex -> u8 I/Q, src/rtlsdr_encode.rs) and FileSink (src/file_sink.rs).  The oracle is pinned on the reference's own
#[test] vectors; the GPU kernel is bit-exact against it; the blocks are driven like the reference's tests; a capture file
goes file -> GPU filters -> file and a decoded-then-encoded byte stream comes back unchanged."""
import numpy as np
import pytest

from oracle import oracle as O

# src/rtlsdr_encode.rs:71-83 `some_input` and :85-94 `clips_to_byte_range` (assert_eq!, exact)
GOLD_IN = np.array([-1.016 - 0.93600005j, -0.85600007 - 0.93600005j, -1.016 - 0.91200006j], np.complex64)
GOLD_OUT = [0, 10, 20, 10, 0, 13]
F32_BYTES = bytes([0, 0, 128, 63, 0, 0, 64, 64, 195, 245, 72, 64, 195, 245, 72, 192])          # src/file_sink.rs:338-343
C32_BYTES = bytes([0, 0, 0, 0, 0, 0, 0, 0, 195, 245, 72, 64, 205, 204, 44, 192])               # :358-361


# ------------------------------------------------------------------ CPU ---
def test_encode_oracle_matches_reference_golden_vectors():
    assert O.rtlsdr_encode(GOLD_IN).tolist() == GOLD_OUT
    assert O.rtlsdr_encode(np.array([-10.0 + 10.0j], np.complex64)).tolist() == [0, 255]
    assert len(O.rtlsdr_encode(np.zeros(0, np.complex64))) == 0


def test_encode_oracle_inverts_decode_for_every_byte_and_handles_specials():
    raw = np.arange(256, dtype=np.uint8).repeat(2)
    assert O.rtlsdr_encode(O.rtlsdr_decode(raw)).tobytes() == raw.tobytes()
    x = np.array([complex(np.nan, np.inf), complex(-np.inf, 0.004), complex(-0.0, 1e30)], np.complex64)
    # NaN -> 0 (saturating cast), +Inf -> 255, -Inf -> 0; 0.004 / 0.008 + 127 = 127.5 -> 128 (half away from zero)
    assert O.rtlsdr_encode(x).tolist() == [0, 255, 0, 128, 127, 255]


def test_encode_plan_is_the_reference_loop():
    import rustradio_b200 as R
    # (in_len, out_free_bytes) -> (consume, produce_bytes, need, wait_on_output), src/rtlsdr_encode.rs:30-51
    assert R.rtlsdr_encode_plan(0, 100) == (0, 0, 1, 0)
    assert R.rtlsdr_encode_plan(3, 100) == (3, 6, 1, 0)
    assert R.rtlsdr_encode_plan(3, 1) == (0, 0, 2, 1)
    assert R.rtlsdr_encode_plan(3, 5) == (2, 4, 2, 1)
    assert R.rtlsdr_encode_plan(3, 6) == (3, 6, 1, 0)


@pytest.fixture(scope="module")
def K():
    from rustradio_b200 import blocks as K
    return K


def test_file_sink_modes_and_bytes_on_host_rings(K, tmp_path):
    """src/file_sink.rs:296-364: fail_create / overwrite / append on an existing file, sink_f32, sink_c32 — host rings need no GPU."""
    from rustradio_b200 import RrcError
    for mode, ok in ((K.FILE_CREATE, False), (K.FILE_OVERWRITE, True), (K.FILE_APPEND, True)):
        w, r = K.new_stream(np.float32, residency=K.HOST)
        if ok:
            K.FileSink(r, "/dev/null", mode)
        else:
            with pytest.raises(RrcError):
                K.FileSink(r, "/dev/null", mode)
            assert len(r) == 0                                       # the caller still owns the stream after a failed build
    for dtype, data, want in ((np.float32, [1.0, 3.0, 3.14, -3.14], F32_BYTES), (np.complex64, [0, 3.14 - 2.7j], C32_BYTES)):
        fn = tmp_path / f"delme_{np.dtype(dtype).name}.bin"
        w, r = K.new_stream(dtype, residency=K.HOST)
        w.write(np.array(data, dtype))
        sink = K.FileSink(r, fn, K.FILE_CREATE, flush=True)
        assert sink.work().kind == K.AGAIN
        ret = sink.work()
        assert ret.kind == K.WAIT and ret.need == 1
        assert fn.read_bytes() == want
        w.write(np.array(data[:1], dtype))                           # append mode keeps what is there
        del sink
        w2, r2 = K.new_stream(dtype, residency=K.HOST)
        w2.write(np.array(data[:1], dtype))
        s2 = K.FileSink(r2, fn, K.FILE_APPEND)
        s2.work()
        assert fn.read_bytes() == want + np.array(data[:1], dtype).tobytes()


# ------------------------------------------------------------------ GPU ---
@pytest.mark.gpu
@pytest.mark.parametrize("n,in_off,out_off", [(0, 0, 0), (1, 0, 0), (2, 0, 0), (7, 0, 0), (100_003, 0, 0), (4096, 1, 0), (4097, 1, 2),
                                              (5000, 0, 1), (5001, 1, 1), (1 << 22, 0, 0)])
def test_encode_kernel_is_bit_exact(n, in_off, out_off):
    """Every alignment class of the fast / byte kernels, specials included, against the oracle."""
    import rustradio_b200 as R
    x = (O.synth_c32(71, 0, n + in_off) * np.float32(1.3)).astype(np.complex64)
    if n + in_off > 12:
        x[in_off + 3] = complex(np.nan, np.inf); x[in_off + 5] = complex(-np.inf, 0.004); x[in_off + 9] = complex(5.0, -5.0)
    din = R.DeviceBuffer.from_numpy(x)
    dout = R.DeviceBuffer(2 * n + out_off + 16)
    R.rtlsdr_encode(din.ptr + 8 * in_off, n, dout.ptr + out_off)
    got = dout.download(np.uint8, 2 * n + out_off)[out_off:]
    assert got.tobytes() == O.rtlsdr_encode(x[in_off:]).tobytes()


@pytest.mark.gpu
def test_encode_host_path_and_decode_roundtrip():
    import rustradio_b200 as R
    raw = O.synth_u8(72, 0, 2 * 3_000_001)
    x = O.rtlsdr_decode(raw)
    assert R.rtlsdr_encode_host(x).tobytes() == raw.tobytes()       # decode -> encode is the identity on bytes


@pytest.mark.gpu
@pytest.mark.parametrize("res", ["DEVICE", "HOST", "HOST_PINNED"])
def test_encode_block_like_the_reference_tests(K, res):
    """src/rtlsdr_encode.rs:57-94: empty / some_input / clips_to_byte_range, plus the WaitForStream(dst, 2) branch."""
    r_ = getattr(K, res)
    src, s1 = K.VectorSource(np.zeros(0, np.complex64), residency=r_)
    assert src.work().kind == K.EOF
    enc, out = K.RtlSdrEncode(s1, residency=r_)
    assert enc.work().kind == K.WAIT and len(out) == 0
    for data, want in ((GOLD_IN, GOLD_OUT), (np.array([-10 + 10j], np.complex64), [0, 255])):
        src, s1 = K.VectorSource(data, residency=r_)
        assert src.work().kind == K.EOF
        enc, out = K.RtlSdrEncode(s1, residency=r_)
        ret = enc.work()
        assert ret.kind == K.WAIT and ret.need == 1
        assert out.read_buf()[0].tolist() == want
    # a full output ring: 2 bytes per sample, WaitForStream(dst, 2) once fewer than 2 bytes are free
    # (device rings round their capacity up to the 2 MiB VMM granularity: feed more than that)
    x = O.synth_c32(73, 0, 1_200_000)
    src, s1 = K.VectorSource(x, size_bytes=16 << 20, residency=r_)
    src.work()
    enc, out = K.RtlSdrEncode(s1, size_bytes=4096, residency=r_)
    cap = out.capacity
    assert cap % 2 == 0 and cap < 2 * len(x)
    ret = enc.work()
    assert ret.kind == K.WAIT and ret.need == 2 and ret.stream_id == out.id and len(out) == cap
    got = []
    for _ in range(10_000):
        got.append(out.read_buf()[0].copy())
        out.consume(len(got[-1]))
        ret = enc.work()
        if ret.need == 1:                                            # WaitForStream(src, 1): input exhausted
            got.append(out.read_buf()[0].copy())
            break
    assert np.concatenate(got).tobytes() == O.rtlsdr_encode(x).tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("res", ["DEVICE", "HOST_PINNED"])
def test_capture_file_to_filtered_file(K, res, tmp_path):
    """u8 capture on disk -> FileSource -> RtlSdrDecode -> FirFilter(/4) -> RtlSdrEncode -> FileSink: the bytes on disk equal
    the oracle chain's (the filter through device rings, codecs bit-exact)."""
    r_ = getattr(K, res)
    raw = O.synth_u8(74, 0, 2 * 200_000)
    fin, fout = tmp_path / "in.u8", tmp_path / "out.u8"
    fin.write_bytes(raw.tobytes())
    taps = O.low_pass_n(1.0, 0.1, 33).astype(np.complex64)
    src, s0 = K.FileSource(fin, np.uint8, size_bytes=1 << 20, residency=r_)
    dec, s1 = K.RtlSdrDecode(s0, size_bytes=1 << 20, residency=K.DEVICE)
    fir, s2 = K.FirFilter(s1, taps, 4, size_bytes=1 << 20, residency=K.DEVICE)
    enc, s3 = K.RtlSdrEncode(s2, size_bytes=1 << 20, residency=r_)
    sink = K.FileSink(s3, fout, K.FILE_CREATE)
    blocks = (src, dec, fir, enc, sink)
    for _ in range(10_000):
        kinds = [b.work().kind for b in blocks]
        if kinds[0] == K.EOF and all(k == K.WAIT for k in kinds[1:]):
            break
    else:
        raise AssertionError("chain did not drain")
    want_c = O.fir(O.rtlsdr_decode(raw), taps, 4)                   # FP32 kernels: sequential f32 order is not bit-pinned ...
    got = np.frombuffer(fout.read_bytes(), np.uint8)
    want = O.rtlsdr_encode(want_c)
    assert len(got) == len(want)
    assert np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 1      # ... so a value on a rounding boundary may move by one code
    assert (got != want).mean() < 1e-3
