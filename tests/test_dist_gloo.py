"""world_size-2 gloo test (CPU) of the N>1 path's host logic: shard assignment, halo arithmetic,
max-over-ranks timing reduction and gathering shards back to the whole-stream result.  The per-shard
compute here is the CPU oracle (this is a test of the sharding logic, not of the kernels); on the GPU box
the same Segment ranges drive the CUDA kernels (tests/test_gpu_parity.py::test_time_segment_shards_equal_whole)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from rustradio_b200 import shard as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 50_000
        x = O.synth_c32(77, 0, n)                                  # every rank can regenerate any window
        taps = O.low_pass_n(1.0, 0.1, 129).astype(np.complex64)
        # --- FIR, time-segment sharded with an ntaps-1 halo
        seg = S.fir_segment(n, len(taps), 3, world, rank)
        y = O.fir(x[seg.in_lo:seg.in_hi], taps, 3)
        assert len(y) == seg.out_hi - seg.out_lo
        # --- FftFilter: shard r filters its inputs + halo, keeps outputs [lo, hi)
        fseg = S.fftfilt_segment(n, len(taps), world, rank)
        full = O.conv_full_f64_fft(x[fseg.in_lo:fseg.in_hi], taps, fseg.in_hi - fseg.in_lo)
        yf = full[fseg.out_lo - fseg.in_lo:].astype(np.complex64)
        # --- resampler: no halo, 64-bit phase only
        rseg = S.resampler_segment(n, 147, 160, world, rank)
        k = np.arange(rseg.out_lo, rseg.out_hi, dtype=np.int64)
        yr = x[rseg.in_lo:rseg.in_hi][(k * 160) // 147 - rseg.in_lo]
        # the same shard through the reference's work() loop started from the shard's counter state
        # (what rrc_resampler_set_state does on the device): identical
        orc = O.Resampler(8, 147, 160)
        orc.set_counter(S.resampler_shard_counter(rseg, 147, 160))
        _, consumed, yr2 = orc.work(x[rseg.in_lo:rseg.in_hi], rseg.out_hi - rseg.out_lo + 3)
        assert consumed == rseg.in_hi - rseg.in_lo and yr2.tobytes() == yr.tobytes()
        # --- config 3 style channel split: rank r owns channels shard_range(nchan, world, r)
        nchan = 5
        c_lo, c_hi = S.shard_range(nchan, world, rank)
        ych = [O.fir(O.synth_c32(78, c * 4000, 4000), taps, 10) for c in range(c_lo, c_hi)]
        # --- gather on rank 0 (object gather: ragged shards)
        gathered = [None] * world
        dist.all_gather_object(gathered, (y, yf, yr, ych))
        # --- max-over-ranks reduction used for timings
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        if rank == 0:
            whole = np.concatenate([g[0] for g in gathered])
            ok_fir = np.array_equal(whole, O.fir(x, taps, 3))
            wf = np.concatenate([g[1] for g in gathered])
            ref = O.conv_full_f64_fft(x, taps, O.fftfilt_out_count(n, len(taps)))
            ok_fft = len(wf) == len(ref) and O.rel_rms(wf, ref) < 1e-6
            wr = np.concatenate([g[2] for g in gathered])
            ok_rs = wr.tobytes() == O.resample(x, 147, 160).tobytes()
            chans = [c for g in gathered for c in g[3]]
            ok_rs = ok_rs and len(chans) == nchan and all(
                np.array_equal(c, O.fir(O.synth_c32(78, i * 4000, 4000), taps, 10)) for i, c in enumerate(chans))
            q.put((ok_fir, ok_fft, ok_rs, float(t.item())))
    finally:
        dist.destroy_process_group()


def test_two_rank_time_segment_sharding_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok_fir, ok_fft, ok_rs, tmax = q.get()
    assert ok_fir and ok_fft and ok_rs
    assert tmax == 2.0


@pytest.mark.parametrize("total,world", [(10, 3), (1024, 8), (5, 8), (0, 2), (239_973, 4)])
def test_shard_range_partitions_exactly(total, world):
    ranges = [S.shard_range(total, world, r) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:]))
    sizes = [hi - lo for lo, hi in ranges]
    assert max(sizes) - min(sizes) <= 1


def test_segments_cover_baseline_configs():
    # config 3: 1024 channels over 8 GPUs -> 128 each
    assert [S.shard_range(1024, 8, r) for r in (0, 7)] == [(0, 128), (896, 1024)]
    # config 4: resampler 2^30 over 8 GPUs, no halo
    segs = [S.resampler_segment(1 << 30, 147, 160, 8, r) for r in range(8)]
    assert segs[0].out_lo == 0 and segs[-1].out_hi == 986_500_301
    assert all(a.out_hi == b.out_lo for a, b in zip(segs[:-1], segs[1:]))
    assert all(s.in_hi - s.in_lo <= (1 << 30) // 8 + 2 for s in segs)
    # config 2: FftFilter 4097 taps, halo = 4096 samples on every shard but the first
    f = [S.fftfilt_segment(1 << 28, 4097, 8, r) for r in range(8)]
    assert f[0].in_lo == 0 and all(s.out_lo - s.in_lo == 4096 for s in f[1:])
    assert f[-1].out_hi == 268_434_089


@pytest.mark.parametrize("interp,deci", [(147, 160), (160, 147), (3, 1), (1, 8), (7, 3)])
def test_resampler_shard_counter_state(interp, deci):
    """Every shard, started from resampler_shard_counter, reproduces its slice of the whole stream through
    the reference's own work() loop; the counter stays in (-interp, 0]."""
    n, world = 10_007, 5
    x = (np.arange(n) * 7919 % 65521).astype(np.uint32)
    whole = O.resample(x, interp, deci)
    parts = []
    for r in range(world):
        seg = S.resampler_segment(n, interp, deci, world, r)
        c0 = S.resampler_shard_counter(seg, interp, deci)
        assert -interp < c0 <= 0
        o = O.Resampler(4, interp, deci)
        o.set_counter(c0)
        _, consumed, y = o.work(x[seg.in_lo:seg.in_hi], seg.out_hi - seg.out_lo)
        assert len(y) == seg.out_hi - seg.out_lo
        parts.append(y)
    assert np.concatenate(parts).tobytes() == whole.tobytes()
