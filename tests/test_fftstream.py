"""FftStream / Fft: the integer part of FftStream::work, the batched forward FFT
kernel against the f64 DFT (rustfft is not vendored: bit pattern unpinned, rel-RMS bar 1e-5), and the
rr::FftStream block driven like src/fft_stream.rs `adds_frame_tags`."""
import numpy as np
import pytest

from oracle import oracle as O


def test_plan_rules():
    import rustradio_b200 as R
    assert R.fftstream_plan(4, 3, 100) == (0, 4, 0)          # src/fft_stream.rs:75-77
    assert R.fftstream_plan(4, 8, 3) == (0, 4, 1)            # :80-82
    assert R.fftstream_plan(4, 8, 100) == (8, 0, 0)
    assert R.fftstream_plan(4, 11, 9) == (8, 0, 0)           # min(in, out) rounded down to frames (:83-84)
    assert R.fftstream_plan(1024, 512_000, 512_000) == (512_000, 0, 0)


def test_create_errors():
    import rustradio_b200 as R
    with pytest.raises(R.RrcError):
        R.Fft(0)                                             # assert_ne!(size, 0) / Err (src/fft.rs:25-27)


@pytest.fixture(scope="module")
def R():
    import rustradio_b200 as R
    assert R.device_count() >= 1
    return R


@pytest.mark.gpu
@pytest.mark.parametrize("size", [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
def test_fft_matches_f64_dft(R, size):
    nframes = max(3, min(700, (1 << 17) // size)) + 1        # odd counts: partial frame groups
    x = O.synth_c32(40 + size % 7, 0, nframes * size)
    got = R.Fft(size).transform(x).reshape(nframes, size)
    want = np.fft.fft(x.astype(np.complex128).reshape(nframes, size), axis=1)
    assert O.rel_rms(got, want) <= 1e-5
    # the reference's own test vector: zeros in, zeros out (src/fft.rs `zeroes`)
    z = R.Fft(size).transform(np.zeros(2 * size, np.complex64))
    assert not z.any()
    # an impulse at n0 is a pure tone across the bins, a tone lands in one bin
    imp = np.zeros(size, np.complex64); imp[size // 3] = 1
    assert O.rel_rms(R.Fft(size).transform(imp), np.fft.fft(imp.astype(np.complex128))) <= 1e-5


@pytest.mark.gpu
def test_fft_in_place_and_host_pipeline(R, monkeypatch):
    size, nframes = 1024, 300
    x = O.synth_c32(44, 0, size * nframes + 77)
    f = R.Fft(size)
    d = R.DeviceBuffer.from_numpy(x)
    f.run(d, nframes, d)                                     # in place
    want = np.fft.fft(x[:size * nframes].astype(np.complex128).reshape(nframes, size), axis=1).ravel()
    assert O.rel_rms(d.download(np.complex64, size * nframes), want) <= 1e-5
    monkeypatch.setenv("RRC_PIPE_CHUNK_LOG2", "14")
    y = f.run_host(x)
    assert len(y) == size * nframes                          # the partial frame is left alone
    assert O.rel_rms(y, want) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("res", ["DEVICE", "HOST"])
def test_block_adds_frame_tags(R, res):
    """src/fft_stream.rs:125-146."""
    from rustradio_b200 import blocks as K
    r = getattr(K, res)
    src, so = K.VectorSource(np.zeros(8, np.complex64), residency=r)
    assert src.work().kind == K.EOF
    fft, fo = K.FftStream(so, 4, residency=r)
    assert fft.work().kind == K.AGAIN
    buf, tags = fo.read_buf()
    assert len(buf) == 8 and not buf.any()
    assert tags == [K.Tag(0, "FftStream::size", ("U64", 4)), K.Tag(0, "FftStream::frame", ("Bool", True)),
                    K.Tag(3, "FftStream::frame", ("Bool", False)), K.Tag(4, "FftStream::size", ("U64", 4)),
                    K.Tag(4, "FftStream::frame", ("Bool", True)), K.Tag(7, "FftStream::frame", ("Bool", False))]
    ret = fft.work()
    assert ret.kind == K.WAIT and ret.need == 4               # WaitForStream(src, size)


@pytest.mark.gpu
def test_block_chain_values_and_partial_frame(R):
    from rustradio_b200 import blocks as K
    size, n = 256, 256 * 37 + 100
    x = O.synth_c32(45, 0, n)
    src, so = K.VectorSource(x)
    fft, fo = K.FftStream(so, size)
    K.graph_run([src, fft])
    got, tags = fo.read_buf()
    assert len(got) == 256 * 37                               # the trailing 100 samples never form a frame
    want = np.fft.fft(x[:256 * 37].astype(np.complex128).reshape(37, size), axis=1).ravel()
    assert O.rel_rms(got, want) <= 1e-5
    assert len(tags) == 3 * 37 and tags[3] == K.Tag(256, "FftStream::size", ("U64", 256))
    with pytest.raises(Exception):
        K.FftStream(K.VectorSource(x)[1], 0)
