"""Host-resident stream ring (C++ mirror of rustradio's circular buffer) against the
restated reference semantics (oracle/blockmodel.py), no GPU needed.
Covers SURVEY Appendix B invariants: contiguous windows across the wrap, tag
re-basing and sorting, consume dropping tags, full/empty windows, EOF by peer drop."""
import numpy as np
import pytest

from oracle import blockmodel as B
from rustradio_b200 import blocks as K

PAGE = 4096


def _tag(pos, key, val):
    return K.Tag(pos, key, val)


def _as_model(tags):
    return [(t.pos, t.key, t.val) for t in tags]


def test_default_stream_capacity_matches_reference():
    w, r = K.new_stream(np.complex64, K.DEFAULT_STREAM_SIZE, K.HOST)
    assert r.capacity == 512_000               # 4_096_000 / 8, src/stream.rs:105
    w2, r2 = K.new_stream(np.float32, K.DEFAULT_STREAM_SIZE, K.HOST)
    assert r2.capacity == 1_024_000
    assert r.id != r2.id and w.id == r.id      # id shared by both ends (src/stream.rs:118)


def test_zero_initialised_and_contiguous_across_wrap():
    w, r = K.new_stream(np.uint32, PAGE, K.HOST)    # 1024 samples
    cap = r.capacity
    assert cap == 1024
    assert w.write(np.arange(1000, dtype=np.uint32)) == 1000
    r.consume(900)
    # 100 used at rpos 900; free window = 924 and wraps the end of the ring
    assert w.free() == 924
    assert w.write(np.arange(5000, 5000 + 924, dtype=np.uint32)) == 924
    got, _ = r.read_buf()
    assert len(got) == 1024
    assert np.array_equal(got, np.concatenate([np.arange(900, 1000), np.arange(5000, 5924)]).astype(np.uint32))
    assert w.write(np.zeros(5, np.uint32)) == 0      # full


def test_random_produce_consume_tags_match_model():
    rng = np.random.default_rng(3)
    w, r = K.new_stream(np.uint32, PAGE, K.HOST)
    m = B.Stream(np.uint32, PAGE)
    counter = 0
    for step in range(600):
        if rng.random() < 0.55:
            n = int(rng.integers(0, 400))
            data = np.arange(counter, counter + n, dtype=np.uint32)
            ntag = int(rng.integers(0, 4))
            tags = [_tag(int(rng.integers(0, max(n, 1))), f"k{step}_{j}", ("U64", step * 10 + j)) for j in range(ntag)]
            wrote = w.write(data, tags)
            mw = m.write_buf()
            assert wrote == min(n, len(mw))
            mw[:wrote] = data[:wrote]
            m.produce(wrote, [B.Tag(t.pos, t.key, t.val) for t in tags if t.pos < wrote])
            counter += wrote
        else:
            got, tags = r.read_buf()
            want, wtags = m.read_buf()
            assert np.array_equal(got, want)
            assert _as_model(tags) == [(t.pos, t.key, t.val) for t in wtags]
            n = int(rng.integers(0, len(got) + 1))
            r.consume(n)
            m.consume(n)
    got, tags = r.read_buf()
    want, wtags = m.read_buf()
    assert np.array_equal(got, want) and _as_model(tags) == [(t.pos, t.key, t.val) for t in wtags]


def test_tag_value_kinds_round_trip():
    w, r = K.new_stream(np.float32, PAGE, K.HOST)
    tags = [_tag(0, "a", ("String", "hello")), _tag(0, "b", ("Float", 1.5)), _tag(1, "c", ("Bool", True)),
            _tag(2, "d", ("U64", 2 ** 63 + 5)), _tag(2, "e", ("I64", -7))]
    w.write(np.zeros(3, np.float32), tags)
    _, got = r.read_buf()
    assert got == tags                        # sorted by pos, stable within a position


def test_eof_when_writer_dropped_and_empty():
    w, r = K.new_stream(np.uint8, PAGE, K.HOST)
    w.write(np.arange(10, dtype=np.uint8))
    assert not r.eof()
    w.drop()
    assert not r.eof()                        # still 10 samples to read (src/stream.rs:237-246)
    r.consume(10)
    assert r.eof()


def test_consume_more_than_available_is_an_error():
    w, r = K.new_stream(np.uint8, PAGE, K.HOST)
    w.write(np.arange(10, dtype=np.uint8))
    with pytest.raises(K.RrcError):
        r.consume(11)


def test_host_stream_size_must_be_page_multiple():
    with pytest.raises(K.RrcError):
        K.new_stream(np.uint8, 1000, K.HOST)
