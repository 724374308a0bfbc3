#!/usr/bin/env python
"""bench.py — throughput of the B200 filtering hot path on BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W [--config c2] [--impl reference]

Default workload (BASELINE.json configs[1], "c2"): FftFilter, 4097-tap low-pass,
2^28 synthetic c32 samples, single stream per GPU.  One "step" = one pass of the
hot path over one such batch.  With N > 1 (torchrun, one rank per GPU) every
rank filters its own independent 2^28-sample capture (weak scaling, no
data-path collective).  Other configs (c1 FIR, c3 channelizer, c4 resampler)
are selectable with --config for the per-kernel numbers in DESIGN.md.

Prints ONE JSON line on rank 0 (contract in the task statement):
  value     = input Msamples/s, inputs resident in HBM (CUDA events, max over ranks)
  e2e       = same metric through the C ABI's *_run_host entry point with
              pinned HOST buffers (H2D + kernel + D2H inside the timed region)
  roofline  = algorithmic bytes of the dominant kernel / its measured duration
              against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline = oracle port (restated CPU reference) timed on this host
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 0x5EED0000


# ----------------------------------------------------------------- configs --
def cfg_c1():
    return dict(name="c1", op="fir", ntaps=64, deci=1, n=1 << 24, cutoff=0.1, dtype="c32", rotate=6,
                desc="FirFilter c32 low-pass 64 taps, deci 1, 2^24 samples")


def cfg_c1f():
    return dict(name="c1f", op="fir", ntaps=64, deci=1, n=1 << 25, cutoff=0.1, dtype="f32",
                desc="FirFilter<Float> f32 low-pass 64 taps, deci 1, 2^25 f32 samples (config 1's bytes as a real stream)")


def cfg_c1d2():
    return dict(name="c1d2", op="fir", ntaps=127, deci=2, n=1 << 24, cutoff=0.2, dtype="c32",
                desc="FirFilter c32 low-pass 127 taps, deci 2, 2^24 samples (half-band decimator shape)")


def cfg_c2():
    return dict(name="c2", op="fftfilt", ntaps=4097, n=1 << 28, cutoff=0.05, dtype="c32",
                desc="FftFilter 4097-tap low-pass, 2^28 c32 samples, single stream per GPU")


def cfg_c3():
    return dict(name="c3", op="fir_demod", ntaps=255, deci=10, nchan=1024, n=2_400_000, dtype="c32",
                desc="rtl_fm channelizer: 1024 ch x 2400000 c32 (1 s @2.4 Msps), 255-tap /10 FIR + QuadratureDemod fused")


def cfg_c4():
    return dict(name="c4", op="resample", interp=147, deci=160, n=1 << 30, dtype="f32",
                desc="RationalResampler 147/160 on 2^30 f32 samples")


def cfg_c5():
    return dict(name="c5", op="fftfilt_decim", ntaps=16385, deci=8, n=1 << 30, cutoff=0.05, dtype="c32",
                desc="wideband: one 2^30-sample c32 capture per GPU, 16385-tap FftFilter + decimate-by-8 fused")


# SURVEY 8f rank 1 (RtlSdrDecode and u8 I/Q ingest): the same shapes fed with RTL-SDR bytes, the
# decode fused into the kernels' first load, and the stand-alone decode block.
def cfg_c3u8():
    c = cfg_c3()
    c.update(name="c3u8", in_u8=True, desc=c["desc"].replace("c32 (", "u8 I/Q (") + ", RtlSdrDecode fused into the tile load")
    return c


def cfg_c5u8():
    c = cfg_c5()
    c.update(name="c5u8", in_u8=True, desc="wideband: one 2^30-sample u8 I/Q capture per GPU, RtlSdrDecode + 16385-tap FftFilter + decimate-by-8 fused")
    return c


def cfg_f1():
    return dict(name="f1", op="decode", n=1 << 29, dtype="c32",
                desc="RtlSdrDecode u8 I/Q -> c32, 2^29 samples (1 GiB in, 4 GiB out)")


def cfg_a12():
    return dict(name="a12", op="fftfilt_real", ntaps=4097, n=1 << 29, cutoff=0.05, dtype="f32",
                desc="FftFilterFloat 4097 real taps, 2^29 f32 samples (real-stream kernel mode: two real blocks per transform)")


def cfg_f2():
    size = int(os.environ.get("RRC_BENCH_FFT_SIZE", "1024"))     # 1024 = the spectrum-display size of the examples
    return dict(name="f2", op="fft", size=size, n=1 << 28, dtype="c32",
                desc=f"FftStream forward FFT, {size}-point frames, 2^28 c32 samples")


# Hilbert and the sample-wise neighbours.
def cfg_h1():
    return dict(name="h1", op="hilbert", ntaps=65, n=1 << 29, dtype="f32",
                desc="Hilbert 65 taps (Hamming), 2^29 f32 samples -> c32 (the examples' ax25/bell202 shape)")


def cfg_e1():
    return dict(name="e1", op="mulconst", n=1 << 29, dtype="c32", desc="MultiplyConst<Complex>, 2^29 c32 samples")


def cfg_e2():
    return dict(name="e2", op="mag2", n=1 << 29, dtype="c32", desc="ComplexToMag2, 2^29 c32 samples -> f32")


def cfg_e3():
    return dict(name="e3", op="tee", n=1 << 29, dtype="c32", desc="Tee<Complex>, 2^29 c32 samples -> two copies")


def cfg_e4():
    return dict(name="e4", op="iqbalance", n=1 << 29, dtype="c32", desc="IqBalance alpha=2.08e-6 (tau 0.2 s @2.4 Msps), 2^29 c32 samples")


CONFIGS = {"c1f": cfg_c1f, "c1d2": cfg_c1d2, "h1": cfg_h1, "e1": cfg_e1, "e2": cfg_e2, "e3": cfg_e3, "e4": cfg_e4, "a12": cfg_a12, "f2": cfg_f2, "c1": cfg_c1, "c2": cfg_c2, "c3": cfg_c3, "c4": cfg_c4, "c5": cfg_c5, "c3u8": cfg_c3u8, "c5u8": cfg_c5u8, "f1": cfg_f1}


def low_pass_taps(ntaps: int, cutoff: float) -> np.ndarray:
    """Hamming (a0 = 25/46) windowed-sinc low-pass of an exact length, cutoff as a fraction of fs
    (the rustradio low_pass formula, src/fir.rs:631-655, for an explicit ntaps).  Host-side, one-off."""
    a0 = np.float32(25.0 / 46.0)
    k = np.arange(ntaps, dtype=np.float32)
    win = a0 - (np.float32(1) - a0) * np.cos(np.float32(2 * np.pi) * k / np.float32(max(ntaps - 1, 1)))
    m = (ntaps - 1) // 2
    nn = (np.arange(ntaps) - m).astype(np.float32)
    w0 = np.float32(2 * np.pi * cutoff)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(nn == 0, w0 / np.float32(np.pi), np.sin(nn * w0) / (nn * np.float32(np.pi))).astype(np.float32) * win
    gain = t[m] + 2 * t[m + 1:m + 1 + m].sum(dtype=np.float32)
    return (t / gain).astype(np.float32)


def taps_for(cfg):
    if cfg["op"] == "fir_demod":
        return low_pass_taps(cfg["ntaps"], 100e3 / 2.4e6).astype(np.complex64)
    t = low_pass_taps(cfg["ntaps"], cfg["cutoff"])
    return t if (cfg["op"] == "fir" and cfg["dtype"] == "f32") else t.astype(np.complex64)


def alg_bytes(cfg, n_in, n_out):
    """algorithmic bytes per step."""
    ib = 2 if cfg.get("in_u8") else 8
    if cfg["op"] == "fir" and cfg["dtype"] == "f32":
        return 4 * n_in + 4 * n_out
    if cfg["op"] in ("fir", "fftfilt", "fftfilt_decim"):
        return ib * n_in + 8 * n_out
    if cfg["op"] == "fir_demod":
        return ib * n_in + 4 * n_out
    if cfg["op"] == "decode":
        return 2 * n_in + 8 * n_out
    if cfg["op"] == "fft":
        return 8 * n_in + 8 * n_out
    if cfg["op"] == "fftfilt_real":
        return 4 * n_in + 4 * n_out
    if cfg["op"] == "resample":
        return 4 * (n_in + n_out)
    if cfg["op"] == "hilbert":
        return 4 * n_in + 8 * n_out
    if cfg["op"] in ("mulconst", "iqbalance"):
        return 8 * n_in + 8 * n_out
    if cfg["op"] == "mag2":
        return 8 * n_in + 4 * n_out
    if cfg["op"] == "tee":
        return 8 * n_in + 16 * n_out
    raise ValueError(cfg["op"])


def alg_flops(cfg, n_in, n_out, fir=None):
    """SURVEY 8(d) flop formulas.  FIR: 8*ntaps per output (4*ntaps when the kernel's declared real-tap
    fast path is active); FFT filter: blocks*(2*5*F*log2 F + 6F + 2*ntaps) at the reference's F."""
    op = cfg["op"]
    if op in ("fir", "fir_demod"):
        per = 2 if cfg["dtype"] == "f32" else 4 if (fir is not None and fir.uses_real_taps) else 8
        nout = n_out + (cfg.get("nchan", 0) if op == "fir_demod" else 0)      # FIR outputs = demod outputs + 1 per channel
        return per * cfg["ntaps"] * nout
    if op in ("fftfilt", "fftfilt_decim", "fftfilt_real"):
        F = 2
        while F < 2 * cfg["ntaps"] - 1:
            F *= 2
        F = max(F, 2)
        f_ref = 1
        while f_ref < cfg["ntaps"]:
            f_ref *= 2
        f_ref *= 2                                                   # calc_fft_size, src/fft_filter.rs:36-42
        blocks = n_in / (f_ref - cfg["ntaps"])
        return blocks * (2 * 5 * f_ref * np.log2(f_ref) + 6 * f_ref + 2 * cfg["ntaps"])
    if op == "fft":
        return (n_in / cfg["size"]) * 5 * cfg["size"] * np.log2(cfg["size"])
    if op == "hilbert":
        return 2 * cfg["ntaps"] * n_out
    return 0


# ---------------------------------------------------------------- clocks ----
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [c.strip() for c in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nme, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------ GPU arm -------
KERNEL_NAMES = {
    "fftfilt": "fftfilt_tmh_kernel<.,.,false,true> (TMA-staged input; spectrum and phase-B twiddles in tensor memory; RRC_FFTFILT_VARIANT=36 selects fftfilt_tma_kernel, 32 the LDG kernel)",
    "fir": "fir_poly_kernel<float2,float,1,false,16>", "fir_demod": "fir_rt_kernel<10,DEMOD,8,2>",
    "resample": "resample_kernel", "decode": "rtlsdr_decode_kernel", "fft": "fftstream_kernel<10>",
    "fftfilt_real": "fftfilt_kernel (real-stream mode)",
    "fftfilt_decim": "fftfilt_poly_kernel<4,false> (polyphase: 8 forward 16384-point transforms + 1 inverse per block, sum in TMEM, 4-CTA cluster; RRC_FFTFILT_NO_POLY=1 selects fftfilt_fold_kernel<4>)",
    "hilbert": "hilbert_half_kernel + history update", "mulconst": "map_kernel<MAP_MUL_C32>", "mag2": "mag2_kernel",
    "tee": "tee_kernel<uint4>",
    "iqbalance": "iq_tile_kernel<false> + iq_carry_kernel + iq_tile_kernel<true> (input read twice: 24 B/sample of traffic vs 16 algorithmic)",
}


class Ctx:
    """Per-process CUDA context of the bench: rank / device / torch stream / torch.distributed."""

    def __init__(self):
        import torch
        import rustradio_b200 as R
        self.torch, self.R = torch, R
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", rank=self.rank, world_size=self.world, device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.dev = self.local
        self.stream = torch.cuda.current_stream().cuda_stream
        self.device = f"cuda:{self.dev}"

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if not self.dist:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if not self.dist:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t)
        return float(t.item())

    def gather_bytes(self, b: bytes) -> list:
        """all_gather of one small bytes object per rank (IPC handles)."""
        if not self.dist:
            return [b]
        out = [None] * self.world
        self.dist.all_gather_object(out, b)
        return out


class Workload:
    """One config's device-resident step: handle, buffers, step(), algorithmic sizes."""
    def __init__(self, **kw):
        self.keep = []           # objects that must outlive the steps (buffers, IPC mappings)
        self.e2e = None          # (fn() -> outputs, h2d bytes, d2h bytes, description) built lazily by make_e2e
        self.scaling = "weak"
        self.parallelism = ""
        self.__dict__.update(kw)


def seed_of(cfg, rank=0):
    return SEED + (int(cfg["name"][1]) if cfg["name"][1].isdigit() else 9) + 1000 * rank


def synth_input(ctx, cfg, seed, nsamp, first=0):
    """Device-resident synthetic input: c32 white noise, or (u8 ingest) the same noise quantised to
    RTL-SDR bytes by an untimed setup step."""
    torch, R = ctx.torch, ctx.R
    if not cfg.get("in_u8"):
        t = torch.empty(2 * nsamp, dtype=torch.float32, device=ctx.device)
        R.synth_f32(t, seed, 2 * first, 2 * nsamp, ctx.dev, ctx.stream)
        return t
    out = torch.empty(2 * nsamp, dtype=torch.uint8, device=ctx.device)
    chunk = 1 << 27
    tmp = torch.empty(min(chunk, 2 * nsamp), dtype=torch.float32, device=ctx.device)
    for o in range(0, 2 * nsamp, chunk):
        m = min(chunk, 2 * nsamp - o)
        R.synth_f32(tmp, seed, 2 * first + o, m, ctx.dev, ctx.stream)
        torch.cuda.synchronize()
        out[o:o + m] = ((tmp[:m] + 1.0) * 128.0).floor_().clamp_(0, 255).to(torch.uint8)
    del tmp
    return out


def build_workload(ctx, cfg, split="capture"):
    """split: 'capture' = every rank its own capture / all channels (weak); 'time' = ONE capture split by
    time segment across the ranks, halo over NVLink peer memory (strong); 'channel' = config 3's channels
    split across the ranks (strong)."""
    torch, R = ctx.torch, ctx.R
    rank, world, dev, stream = ctx.rank, ctx.world, ctx.dev, ctx.stream
    op = cfg["op"]
    u8 = bool(cfg.get("in_u8"))
    seed = seed_of(cfg, rank if split == "capture" else 0)
    W = Workload(cfg=cfg, op=op, launches_per_step=1)
    dev_f32 = lambda n: torch.empty(n, dtype=torch.float32, device=ctx.device)

    if op == "fftfilt" and split == "time":
        # ONE capture split by time segment: rank r filters outputs [lo, hi); its ntaps-1 sample
        # left halo is the tail of rank r-1's input buffer, which the kernel's first block reads DIRECTLY
        # through a CUDA-IPC mapping of that buffer (NVLink loads) -- no collective, no copy, no host wait.
        from rustradio_b200 import shard as S
        n, T1 = cfg["n"], cfg["ntaps"] - 1
        f = R.FftFilt(taps_for(cfg), device=dev)
        seg = S.fftfilt_segment(n, cfg["ntaps"], world, rank)
        m = seg.out_hi - seg.out_lo
        din = R.DeviceBuffer(m * 8, dev)                       # plain cudaMalloc: exportable
        dout = dev_f32(2 * m)
        R.synth_f32(din, seed, 2 * seg.out_lo, 2 * m, dev, stream)
        handles = ctx.gather_bytes(R.ipc_export(din))
        sizes = ctx.gather_bytes(m.to_bytes(8, "little"))
        if rank > 0:
            peer = R.IpcMapping(handles[rank - 1], dev)
            W.keep.append(peer)
            halo_ptr = peer.ptr + (int.from_bytes(sizes[rank - 1], "little") - T1) * 8
        else:
            zeros = torch.zeros(2 * T1, dtype=torch.float32, device=ctx.device)    # the stream's initial state (src/fft_filter.rs:270)
            W.keep.append(zeros)
            halo_ptr = zeros.data_ptr()
        ctx.barrier()

        def step():
            f.set_history_ptr(halo_ptr, T1)                    # no copy, no memset: the first block reads through the pointer
            f.run(din, m, dout, stream)
        W.__dict__.update(f=f, step=step, n_in=m, n_out=m, units=m, launches_per_step=1, scaling="strong",
                          parallelism=f"one {n}-sample capture, time-segment sharded x{world}; (ntaps-1)-sample halo read by the kernel "
                                      "through a CUDA-IPC mapping of the left neighbour's buffer (NVLink), no collective")
        W.keep += [din, dout]
    elif op == "fftfilt":
        n = cfg["n"]
        f = R.FftFilt(taps_for(cfg), device=dev)
        n_in = n_out = (n // f.nsamples) * f.nsamples            # reference count rule (whole blocks)
        din = synth_input(ctx, cfg, seed, n)
        dout = dev_f32(2 * n_out)

        def step():
            f.run(din, n_in, dout, stream)
        W.__dict__.update(f=f, step=step, n_in=n_in, n_out=n_out, units=n_in, launches_per_step=2, din=din, n_host=n)
        W.keep += [din, dout]
    elif op == "fftfilt_real":
        n = cfg["n"]
        f = R.FftFilt(taps_for(cfg).real.astype(np.float32), device=dev, real=True)
        n_in = n_out = (n // f.nsamples) * f.nsamples
        din = dev_f32(n)
        dout = dev_f32(n_out)
        R.synth_f32(din, seed, 0, n, dev, stream)

        def step():
            f.run(din, n_in, dout, stream)
        W.__dict__.update(f=f, step=step, n_in=n_in, n_out=n_out, units=n_in, launches_per_step=2, din=din, n_host=n)
        W.keep += [din, dout]
    elif op == "fftfilt_decim":
        n = cfg["n"]
        f = R.FftFilt(taps_for(cfg), device=dev)
        n_in = (n // f.nsamples) * f.nsamples
        n_out = (n_in + cfg["deci"] - 1) // cfg["deci"]
        if u8:
            f.set_input_u8iq(True)
        din = synth_input(ctx, cfg, seed, n)
        dout = dev_f32(2 * n_out)

        def step():
            assert f.decim_run(din, n_in, cfg["deci"], 0, dout, stream) == n_out
        W.__dict__.update(f=f, step=step, n_in=n_in, n_out=n_out, units=n_in, launches_per_step=3, din=din, n_host=n)
        W.keep += [din, dout]
    elif op == "fir":
        n = cfg["n"]
        f = R.Fir(taps_for(cfg), deci=cfg["deci"], device=dev)
        n_out = f.out_count(n)
        fl = 1 if cfg["dtype"] == "f32" else 2                # floats per sample
        # the input (128 MiB for config 1) is about the size of the L2: rotate over NBUF distinct input buffers
        # so that no step finds its input in L2 (HBM numbers, not L2 numbers)
        nbuf = int(cfg.get("rotate", 1))
        dins = []
        for b in range(nbuf):
            t = dev_f32(fl * n)
            R.synth_f32(t, seed + 17 * b, 0, fl * n, dev, stream)
            dins.append(t)
        # ... and over NBUF distinct output buffers, so that a step's stores are not absorbed by lines the previous
        # step left dirty in L2 at the same addresses: in the steady state every output byte is written back to HBM
        douts = [dev_f32(fl * n_out) for _ in range(nbuf)]
        dout = douts[0]
        state = {"i": 0}

        def step():
            f.run(dins[state["i"] % nbuf], n, douts[state["i"] % nbuf], n_out, stream)
            state["i"] += 1
        W.__dict__.update(f=f, step=step, n_in=n, n_out=n_out, units=n, din=dins[0], n_host=n)
        W.keep += dins + douts
    elif op == "fir_demod":
        from rustradio_b200 import shard as S
        n = cfg["n"]
        c_lo, c_hi = S.shard_range(cfg["nchan"], world, rank) if split == "channel" else (0, cfg["nchan"])
        nchan = c_hi - c_lo
        f = R.Fir(taps_for(cfg), deci=cfg["deci"], device=dev)
        out_n = f.out_count(n)
        need = (out_n - 1) * cfg["deci"] + cfg["ntaps"]
        if u8:
            f.set_input_u8iq(True)
        din = synth_input(ctx, cfg, seed, n * nchan, first=c_lo * n)
        dout = dev_f32((out_n - 1) * nchan)

        def step():
            f.demod_run_batch(din, n, need, 1.0, dout, out_n - 1, out_n, nchan, stream)
        W.__dict__.update(f=f, step=step, n_in=n * nchan, n_out=(out_n - 1) * nchan, units=n * nchan, din=din, nchan=nchan,
                          n_host=n * nchan)
        if split == "channel":
            W.scaling = "strong"
            W.parallelism = f"{cfg['nchan']} channels split by channel x{world} (rank r owns channels shard_range(1024, W, r)), no inter-GPU traffic"
        W.keep += [din, dout]
    elif op == "decode":
        n = cfg["n"]
        c2 = dict(cfg, in_u8=True)
        din = synth_input(ctx, c2, seed, n)
        dout = dev_f32(2 * n)

        def step():
            R.rtlsdr_decode(din, 2 * n, dout, dev, stream)
        W.__dict__.update(f=None, step=step, n_in=n, n_out=n, units=n)
        W.keep += [din, dout]
    elif op == "fft":
        n = cfg["n"]
        f = R.Fft(cfg["size"], device=dev)
        din = synth_input(ctx, cfg, seed, n)
        dout = dev_f32(2 * n)

        def step():
            f.run(din, n // cfg["size"], dout, stream)
        W.__dict__.update(f=f, step=step, n_in=n, n_out=n, units=n, din=din, n_host=n)
        W.keep += [din, dout]
    elif op in ("hilbert", "mulconst", "mag2", "tee", "iqbalance"):
        n = cfg["n"]
        fin = 1 if op == "hilbert" else 2                       # floats per input sample
        fout = 1 if op == "mag2" else 2
        din = dev_f32(fin * n)
        dout = dev_f32(fout * n)
        R.synth_f32(din, seed, 0, fin * n, dev, stream)
        L = R.lib()
        lps, f = 1, None
        if op == "hilbert":
            f = R.Hilbert(cfg["ntaps"], device=dev)
            lps = 2
            step = lambda: f.run(din, n, dout, stream)
        elif op == "iqbalance":
            f = R.IqBalance(R.iq_balance_alpha_from_tau(2_400_000, 0.2), device=dev)
            lps = 3
            step = lambda: f.run(din, n, dout, stream)
        elif op == "mulconst":
            def step():
                assert L.rrc_multiply_const_c32_run(dev, din.data_ptr(), n, 0.3, -1.7, dout.data_ptr(), stream) == 0
        elif op == "mag2":
            def step():
                assert L.rrc_complex_to_mag2_run(dev, din.data_ptr(), n, dout.data_ptr(), stream) == 0
        else:
            dout2 = dev_f32(2 * n)
            W.keep.append(dout2)

            def step():
                assert L.rrc_tee_run(dev, din.data_ptr(), 8 * n, dout.data_ptr(), dout2.data_ptr(), stream) == 0
        W.__dict__.update(f=f, step=step, n_in=n, n_out=n, units=n, launches_per_step=lps)
        W.keep += [din, dout]
    elif op == "resample":
        from rustradio_b200 import shard as S
        n, I, D = cfg["n"], cfg["interp"], cfg["deci"]
        f = R.Resampler(4, I, D, device=dev)
        if split == "time":
            # time-segment shard (no halo): rank r owns outputs [k_lo, k_hi) and starts mid-stream at input
            # s = floor(k_lo*D/I) with the reference's counter state s*I - k_lo*D (rrc_resampler_set_state)
            seg = S.resampler_segment(n, I, D, world, rank)
            n_in, n_out, c0 = seg.in_hi - seg.in_lo, seg.out_hi - seg.out_lo, seg.in_lo * I - seg.out_lo * D
            first = seg.in_lo
            W.scaling = "strong"
            W.parallelism = f"one {n}-sample stream, time-segment sharded x{world} (no halo; counter state seeded per shard), no inter-GPU traffic"
        else:
            n_in, n_out, c0, first = n, (n * I + D - 1) // D, 0, 0
        din = dev_f32(n_in)
        dout = dev_f32(n_out + 16)
        R.synth_f32(din, seed, first, n_in, dev, stream)

        def step():
            f.set_state(c0)
            c, p, w = f.run(din, n_in, dout, n_out + 16, stream)
            assert (c, p) == (n_in, n_out)
        W.__dict__.update(f=f, step=step, n_in=n_in, n_out=n_out, units=n_in, din=din, n_host=n_in)
        W.keep += [din, dout]
    else:
        raise SystemExit(f"unknown op {op}")
    if not W.parallelism:
        W.parallelism = f"independent stream per GPU x{world}"
    return W


def time_steps(ctx, step, steps, warmup):
    """W warm-ups, then exactly `steps` steps between CUDA events on the launching stream, bracketed by
    barrier + synchronize; returns (max-over-ranks ms per step, launches counted on this rank)."""
    torch, R = ctx.torch, ctx.R
    for _ in range(max(warmup, 3)):
        step()
    ctx.barrier()
    l0 = R.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(ev0.elapsed_time(ev1))
    return ms / steps, R.launch_count() - l0


def job_value(ctx, W, ms_per_step):
    """Whole-job input Msamples/s: units of all ranks / max-over-ranks time."""
    total = ctx.sum_over_ranks(float(W.units))
    return total / (ms_per_step * 1e-3) / 1e6


class HostPool:
    """One pinned input and one pinned output region reused by every config's end-to-end leg (page-locking
    gigabytes costs seconds; the pages sit on the GPU's NUMA node when the platform exposes one)."""

    def __init__(self, ctx):
        self.ctx, self.inb, self.outb = ctx, None, None

    def get(self, in_bytes, out_bytes):
        R = self.ctx.R
        if self.inb is None or self.inb.count < in_bytes:
            if self.inb:
                self.inb.free()
            self.inb = R.PinnedBuffer(np.uint8, in_bytes, near_device=self.ctx.dev)
        if self.outb is None or self.outb.count < out_bytes:
            if self.outb:
                self.outb.free()
            self.outb = R.PinnedBuffer(np.uint8, out_bytes, near_device=self.ctx.dev)
        return self.inb, self.outb

    def free(self):
        for b in (self.inb, self.outb):
            if b:
                b.free()
        self.inb = self.outb = None


def measure_e2e(ctx, W, pool, steps):
    """The same metric through the C ABI's *_run_host entry point with pinned HOST buffers: every step copies
    the step's whole input host->device and the whole result device->host inside the timed region
    (chunked, double buffered, three streams: csrc/pipeline.cuh)."""
    torch, R = ctx.torch, ctx.R
    cfg, op, f = W.cfg, W.op, W.f
    if op not in ("fftfilt", "fir", "fftfilt_decim", "fft", "fftfilt_real", "fir_demod", "resample") or W.scaling != "weak":
        return None
    u8 = bool(cfg.get("in_u8"))
    real = op in ("fftfilt_real", "resample") or (op == "fir" and cfg["dtype"] == "f32")
    ib = 2 if u8 else 4 if real else 8
    ob = 4 if (real or op == "fir_demod") else 8
    n_host = W.n_host
    out_elems = W.n_out + (16 if op == "resample" else 0)
    hin, hout = pool.get(n_host * ib, out_elems * ob)
    R.lib().rrc_memcpy_d2h(ctx.dev, hin.ptr, W.din.data_ptr(), n_host * ib, ctx.stream)
    torch.cuda.synchronize()
    in_dt = np.uint8 if u8 else np.float32 if real else np.complex64
    out_dt = np.float32 if ob == 4 else np.complex64
    xin = hin.array[: n_host * ib].view(in_dt)
    xout = hout.array[: out_elems * ob].view(out_dt)
    stateful = op in ("fftfilt", "fftfilt_decim", "fftfilt_real")
    if op == "fftfilt_decim":
        call = lambda: len(f.decim_run_host(xin, cfg["deci"], xout))
    elif op == "fir_demod":
        call = lambda: f.demod_run_host_batch(xin, cfg["n"], W.nchan, 1.0, xout[: W.n_out]) * W.nchan
    elif op == "resample":
        def call():
            f.set_state(0)
            c, p = f.run_host_into(xin, xout)
            assert (c, p) == (W.n_in, W.n_out)
            return p
    else:
        call = lambda: len(f.run_host(xin, xout))
    for _ in range(1):
        if stateful:
            f.reset()
        call()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        if stateful:
            f.reset()
        got = call()
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    ctx.barrier()
    total = ctx.sum_over_ranks(float(W.units))
    h2d, d2h = int(ib * (n_host if op in ("fir", "fir_demod", "resample") else W.n_in)), int(ob * got)
    return {"value": total / (dt / steps) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": steps, "ms_per_step": dt / steps * 1e3,
            "pcie_gbs_per_gpu": {"h2d": h2d / (dt / steps) / 1e9, "d2h": d2h / (dt / steps) / 1e9},
            "input": "u8 I/Q bytes (RtlSdrDecode fused into the first load)" if u8 else ("f32" if real else "c32"),
            "timer": "host wall clock around rrc_*_run_host (returns after the last D2H completes), max over ranks"}


def copy_ceiling(ctx, pool, h2d_bytes, d2h_bytes, steps=3):
    """The e2e leg's roofline: the same bytes moved with bare cudaMemcpyAsync (H2D and D2H concurrently on
    two streams, pinned memory near the GPU), no kernel.  Per rank, max over ranks -> node aggregate."""
    torch, R = ctx.torch, ctx.R
    L = R.lib()
    hin, hout = pool.get(h2d_bytes, d2h_bytes)
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=ctx.device)
    d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=ctx.device)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def go():
        L.rrc_memcpy_h2d(ctx.dev, d_in.data_ptr(), hin.ptr, h2d_bytes, s1.cuda_stream)
        L.rrc_memcpy_d2h(ctx.dev, hout.ptr, d_out.data_ptr(), d2h_bytes, s2.cuda_stream)
    go()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        go()
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0) / steps
    ctx.barrier()
    del d_in, d_out
    return {"ms_per_step": dt * 1e3, "h2d_gbs_per_gpu": h2d_bytes / dt / 1e9, "d2h_gbs_per_gpu": d2h_bytes / dt / 1e9,
            "node_total_gbs": (h2d_bytes + d2h_bytes) * ctx.world / dt / 1e9,
            "how": "bare cudaMemcpyAsync of the same byte counts, H2D and D2H concurrently, pinned host memory, no kernel"}


def roofline_of(ctx, W, ms_per_step):
    cfg, op, f = W.cfg, W.op, W.f
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = json.loads(peaks_path.read_text())["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ab = alg_bytes(cfg, W.n_in, W.n_out)
    achieved = ab / (ms_per_step * 1e-3) / 1e9
    traffic_path = ROOT / "profiles" / "traffic.json"
    traffic = json.loads(traffic_path.read_text()).get(cfg["name"]) if traffic_path.exists() else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": ab,
                "kernel": KERNEL_NAMES[op], "duration_ms": ms_per_step,
                "note": "duration = CUDA-event time of the whole step on the launching stream / steps; the step is this one kernel"
                        + (" plus a <3 us history-update kernel" if op == "fftfilt" else "")}
    if op in ("fir", "fir_demod"):
        roofline["kernel"] = f.kernel_name + (" + fused QuadratureDemod epilogue" if op == "fir_demod" else "")
    if op in ("fir", "fir_demod") and f.uses_tensor_cores:
        # Declared: the real-tap c32 FIR runs as a block-scaled fp16x3 Toeplitz product on the tensor cores (fir_tc.cuh).
        nout_fir = W.n_out + (getattr(W, "nchan", 0) if op == "fir_demod" else 0)
        if "fir_tc5_kernel" in roofline["kernel"]:
            # tcgen05 path (fir_tc5.cu): 6 MMAs (M128 N64 K16) per k-step per 8192-output tile, K = 127 + ntaps padded to 16
            ks = (127 + cfg["ntaps"] + 15) // 16
            mmas = 6 * ks * nout_fir / 8192
            mma_peak = 148 * 1.965e9 / 32                             # tensor-pipe floor of the shape: 128 * 64 / 256 cycles per MMA
            roofline["tensor"] = {"tcgen05_mma_m128n64k16_per_launch": mmas, "achieved_mma_per_s": mmas / (ms_per_step * 1e-3),
                                  "peak_mma_per_s": mma_peak, "frac": mmas / (ms_per_step * 1e-3) / mma_peak,
                                  "peak_source": "tcgen05 floor max(M,128)*N/256 = 32 cycles per MMA (measured 32.4 in the kernel's trace)"}
        else:
            roofline["kernel"] += " [block-scaled fp16x3 Toeplitz product, mma.m16n8k16 + ldmatrix]"
            ks = (7 * cfg["deci"] + cfg["ntaps"] + 15) // 16         # k-steps of 16 at 8 outputs per block-row (lower bound)
            mmas = 3 * ks * nout_fir / 64                             # three m16n8k16 per k-step per 64 complex outputs
            mma_peak = 148 * 0.46 * 1.965e9                           # measured, profiles/r01_microbench_hmma_rate.txt
            roofline["tensor"] = {"mma_m16n8k16_per_launch": mmas, "achieved_mma_per_s": mmas / (ms_per_step * 1e-3),
                                  "peak_mma_per_s": mma_peak, "frac": mmas / (ms_per_step * 1e-3) / mma_peak,
                                  "peak_source": "measured mma.sync m16n8k16 issue rate, 0.46 per clk per SM (tools/microbench/hmma_rate.cu)"}
    # FP32 side of the roofline: algorithmic flops of the reference formulation against the
    # FP32 FMA rate MEASURED on this pool's B200 (tools/microbench/fp32_pipes.cu: 125 lanes/clk/SM).
    if op == "fir_demod":
        c2 = dict(cfg, nchan=getattr(W, "nchan", cfg["nchan"]))
        flops = alg_flops(c2, W.n_in, W.n_out, f)
    else:
        flops = alg_flops(cfg, W.n_in, W.n_out, f if op == "fir" else None)
    if flops:
        fp_peak = 148 * 125.0 * 2 * 1.965e9 / 1e12
        roofline["fp32"] = {"algorithmic_flops_per_launch": flops, "achieved_tflops": flops / (ms_per_step * 1e-3) / 1e12,
                            "peak_tflops": fp_peak, "frac": flops / (ms_per_step * 1e-3) / 1e12 / fp_peak,
                            "peak_source": "measured FFMA issue rate (profiles/r01_microbench_fp32_pipes.txt) x 2 flop"}
    return roofline


def config_block(cfg, W, world):
    ab = alg_bytes(cfg, W.n_in, W.n_out)
    return {"workload": cfg["desc"], "name": cfg["name"], "samples_per_gpu_per_step": int(W.units),
            "outputs_per_gpu_per_step": int(W.n_out), "parallelism": W.parallelism,
            "l2_policy": ("inputs larger than L2 (>= 0.5 GiB per step vs 126 MB L2)" if ab > 4e8 else
                          f"input 128 MiB ~ L2 size: the steps rotate over {cfg.get('rotate', 1)} distinct input and output buffers ({cfg.get('rotate', 1) * 128} MiB each way) so no step finds its input in L2 or overwrites lines it left dirty there")}


def run_sub(ctx, name, split, args, pool, want_e2e=True, want_cpu=True, steps=None):
    """One config as a sub-record of the default line: device-resident value + roofline (+ e2e, cpu)."""
    cfg = CONFIGS[name]()
    torch = ctx.torch
    rec = {"name": name, "workload": cfg["desc"]}
    W = None
    try:
        W = build_workload(ctx, cfg, split)
        ms, launches = time_steps(ctx, W.step, steps or args.steps, args.warmup)
        rec.update({"value": job_value(ctx, W, ms), "unit": "Msamples/s", "ms_per_step": ms, "scaling": W.scaling,
                    "n_gpus": ctx.world, "gpu_launches": int(launches), "parallelism": W.parallelism,
                    "samples_per_gpu_per_step": int(W.units), "outputs_per_gpu_per_step": int(W.n_out),
                    "roofline": roofline_of(ctx, W, ms)})
        if want_e2e:
            rec["e2e"] = measure_e2e(ctx, W, pool, max(1, min(args.e2e_steps, 2)))
        if want_cpu and ctx.rank == 0 and ctx.world == 1 and not args.no_cpu:
            rec["cpu_baseline"] = cpu_baseline(cfg, threads=1, budget_s=args.sub_cpu_budget)
            rec["cpu_baseline_all_cores"] = cpu_baseline(cfg, threads=os.cpu_count() or 1, budget_s=args.sub_cpu_budget)
    except Exception as e:                                       # a failing sub-record must not take the headline down
        rec["error"] = f"{type(e).__name__}: {e}"
    finally:
        if W is not None:
            W.keep.clear()
            W.__dict__.pop("din", None)
            W.__dict__.pop("step", None)
        del W
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    return rec


def run_gpu(args):
    ctx = Ctx()
    torch, R = ctx.torch, ctx.R
    rank, world = ctx.rank, ctx.world
    cfg = CONFIGS[args.config]()
    if args.n:
        cfg["n"] = args.n
        cfg["desc"] += f" [EXPERIMENT: n overridden to {args.n}]"
    pool = HostPool(ctx)
    split = {"capture": "capture", "time": "time", "channel": "channel"}[args.shard] if world > 1 else "capture"
    W = build_workload(ctx, cfg, split)

    sampler = ClockSampler(ctx.dev)
    if rank == 0:
        sampler.start()
    ms_per_step, launches = time_steps(ctx, W.step, args.steps, args.warmup)
    value = job_value(ctx, W, ms_per_step)
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(ctx, W, pool, max(1, min(args.steps, args.e2e_steps)))
        if e2e is not None and not args.no_ceiling:
            e2e["copy_ceiling"] = copy_ceiling(ctx, pool, e2e["h2d_bytes_per_step"], e2e["d2h_bytes_per_step"])
            e2e["frac_of_copy_ceiling"] = e2e["copy_ceiling"]["ms_per_step"] / e2e["ms_per_step"]
    clocks = sampler.stop() if rank == 0 else None

    # sustained: the headline step back to back for >= args.sustain seconds with its own clock samples
    sustained = None
    if args.sustain > 0:
        s2 = ClockSampler(ctx.dev)
        if rank == 0:
            s2.start()
        nrep = max(args.steps, int(args.sustain * 1e3 / ms_per_step) + 1)
        ms_s, _ = time_steps(ctx, W.step, nrep, 3)
        sustained = {"steps": nrep, "ms_per_step": ms_s, "value": job_value(ctx, W, ms_s), "seconds": nrep * ms_s * 1e-3,
                     "clocks": s2.stop() if rank == 0 else None}
    roofline = roofline_of(ctx, W, ms_per_step)
    cfg_block = config_block(cfg, W, world)
    dtype = "f32" if cfg["dtype"] == "f32" else "c32 (complex f32)"
    scaling = W.scaling
    numa = None
    try:
        numa = {"gpu_numa_node": R.device_numa_node(ctx.dev),
                "nodes_online": Path("/sys/devices/system/node/online").read_text().strip() if Path("/sys/devices/system/node/online").exists() else None}
    except Exception:
        pass
    W.keep.clear(); W.__dict__.pop("din", None); W.__dict__.pop("step", None)
    del W
    import gc
    gc.collect(); torch.cuda.empty_cache()

    # the other BASELINE configs as sub-records (device-resident value, roofline, e2e, CPU baseline), and at
    # N > 1 the north_star's own splits: config 2 / 4 by time segment, config 3 by channel, config 5 by capture
    configs, splits = {}, {}
    if args.config == "c2" and not args.headline_only:
        for name in ("c1", "c3", "c4", "c5"):
            if world == 1:
                configs[name] = run_sub(ctx, name, "capture", args, pool)
        if world == 1:
            for name in ("c3u8", "c5u8"):                     # the PCIe-lean end-to-end forms (2 B/sample in)
                configs[name] = run_sub(ctx, name, "capture", args, pool, want_cpu=False)
        if world > 1:
            splits["c2_time"] = run_sub(ctx, "c2", "time", args, pool, want_e2e=False)
            splits["c3_channel"] = run_sub(ctx, "c3", "channel", args, pool, want_e2e=False)
            splits["c4_time"] = run_sub(ctx, "c4", "time", args, pool, want_e2e=False)
            splits["c5_capture"] = run_sub(ctx, "c5", "capture", args, pool, want_e2e=False)
            splits["c5u8_capture_e2e"] = run_sub(ctx, "c5u8", "capture", args, pool, want_e2e=True)
    pool.free()

    if rank != 0:
        if ctx.dist:
            ctx.dist.destroy_process_group()
        return
    cpu = cpu_all = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_baseline(cfg, threads=1, budget_s=args.cpu_budget)
        cpu_all = cpu_baseline(cfg, threads=os.cpu_count() or 1, budget_s=args.cpu_budget)
    line = {
        "metric": "Msamples/s (c32) FIR/FftFilter/resampler at 1/2/4/8 B200; % of roofline",
        "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": dtype, "data": "synthetic", "config": cfg_block,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "cpu_baseline_all_cores": cpu_all, "sustained": sustained, "numa": numa,
        "timer": "torch.cuda.Event on the launching stream, max over ranks",
    }
    if configs:
        line["configs"] = configs
    if splits:
        line["splits"] = splits
    print(json.dumps(line), flush=True)
    if ctx.dist:
        ctx.dist.destroy_process_group()


# ------------------------------------------------------ CPU reference arm ---
def cpu_baseline(cfg, threads: int, budget_s: float):
    """Times the oracle port (oracle/rr_oracle.c, -O3 AVX2 build, no FMA contraction like rustc)
    on a bounded sample of the same workload.  This is the ONLY place bench.py executes oracle/."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O
    op = cfg["op"]
    taps = taps_for(cfg) if op not in ("resample", "decode", "fft", "hilbert", "mulconst", "mag2", "tee", "iqbalance") else None
    if op == "fftfilt_real":
        # the reference's FftFilterFloat: widen to Complex, complex FftFilter, keep .re (src/fft_filter.rs:428-470)
        per = 1 << 21
        xr = O.synth_f32(SEED + 8, 0, per)
        objs = [O.FftFilt(taps, fast=True) for _ in range(threads)]
        fn = lambda i: len(np.ascontiguousarray(objs[i].run(xr.astype(np.complex64)).real))
        sample = f"{threads} x 2^21 f32 samples per repetition, widen -> overlap-add FftFilter (F=16384) -> .re like the reference"
    elif op == "fftfilt":
        per = 1 << 21
        x = O.synth_c32(SEED + 2, 0, per)
        objs = [O.FftFilt(taps, fast=True) for _ in range(threads)]
        fn = lambda i: len(objs[i].run(x))
        sample = f"{threads} x 2^21 c32 samples per repetition, overlap-add with F=16384 like the reference"
    elif op == "fir":
        per = 1 << 19
        x = O.synth_f32(SEED + 1, 0, per) if cfg["dtype"] == "f32" else O.synth_c32(SEED + 1, 0, per)
        fn = lambda i: len(O.fir(x, taps, cfg["deci"], fast=True))
        sample = f"{threads} x 2^19 {cfg['dtype']} samples per repetition"
    elif op == "fir_demod":
        per = 240_000
        if cfg.get("in_u8"):
            raw = O.synth_u8(SEED + 3, 0, 2 * per)
            fn = lambda i: len(O.quad_demod(O.fir(O.rtlsdr_decode(raw), taps, cfg["deci"], fast=True), fast=True))
            sample = f"{threads} channels x 240000 u8 I/Q samples per repetition (RtlSdrDecode -> FirFilter -> QuadratureDemod)"
        else:
            x = O.synth_c32(SEED + 3, 0, per)
            fn = lambda i: len(O.quad_demod(O.fir(x, taps, cfg["deci"], fast=True), fast=True))
            sample = f"{threads} channels x 240000 c32 samples per repetition"
    elif op == "fftfilt_decim":
        per = 1 << 21
        x = O.synth_u8(SEED + 5, 0, 2 * per) if cfg.get("in_u8") else O.synth_c32(SEED + 5, 0, per)
        objs = [O.FftFilt(taps, fast=True) for _ in range(threads)]
        dec = (lambda v: O.rtlsdr_decode(v)) if cfg.get("in_u8") else (lambda v: v)
        fn = lambda i: len(O.resample(objs[i].run(dec(x)), 1, cfg["deci"]))
        sample = f"{threads} x 2^21 samples per repetition, FftFilter overlap-add with F=65536 like the reference, then RationalResampler(1,8)"
    elif op == "fft":
        per = 1 << 20
        x = O.synth_c32(SEED + 7, 0, per)
        sz = cfg["size"]

        def fn(i):
            y = x.copy()
            for o in range(0, per, sz):
                O.lib(True).orc_fft_c32(y[o:o + sz].ctypes.data, sz, 0)
            return len(y)
        sample = f"{threads} x 2^20 c32 samples per repetition ({sz}-point frames, the oracle's radix-4 FFT)"
    elif op == "hilbert":
        per = 1 << 21
        xr = O.synth_f32(SEED + 9, 0, per)
        objs = [O.Hilbert(cfg["ntaps"]) for _ in range(threads)]
        fn = lambda i: len(objs[i].work(xr))
        sample = f"{threads} x 2^21 f32 samples per repetition (faithful -O2 build: the scalar Fir::filter loop)"
    elif op in ("mulconst", "mag2", "tee", "iqbalance"):
        per = 1 << 23
        x = O.synth_c32(SEED + 9, 0, per)
        iq = [O.IqBalance(2.0833e-6) for _ in range(threads)]
        fn = {"mulconst": lambda i: len(O.multiply_const(x, 0.3 - 1.7j)), "mag2": lambda i: len(O.complex_to_mag2(x)),
              "tee": lambda i: len(x.copy()) + len(x.copy()), "iqbalance": lambda i: len(iq[i].work(x))}[op]
        sample = f"{threads} x 2^23 c32 samples per repetition"
    elif op == "decode":
        per = 1 << 24
        raw = O.synth_u8(SEED + 6, 0, 2 * per)
        fn = lambda i: len(O.rtlsdr_decode(raw))
        sample = f"{threads} x 2^24 samples per repetition"
    else:
        per = 1 << 24
        x = O.synth_f32(SEED + 4, 0, per)
        fn = lambda i: len(O.resample(x, cfg["interp"], cfg["deci"]))
        sample = f"{threads} x 2^24 f32 samples per repetition"
    fn(0)  # warm-up
    reps, t0 = 0, time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        while True:
            list(ex.map(fn, range(threads)))
            reps += 1
            if time.perf_counter() - t0 >= budget_s or reps >= 4096:
                break
    dt = time.perf_counter() - t0
    return {"value": per * threads * reps / dt / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
            "sample": f"{sample}, {reps} repetitions, {dt:.1f} s",
            "note": "restated CPU baseline (oracle port), not rustradio itself: no Rust toolchain in the image; "
                    "the port's scalar radix-4 FFT is slower than rustfft's AVX planner"}


def static_sizes(cfg):
    """(n_in, n_out) per GPU per step from the reference's count rules alone (no device needed)."""
    op = cfg["op"]
    if op in ("fftfilt", "fftfilt_real", "fftfilt_decim"):
        f = 1
        while f < cfg["ntaps"]:
            f <<= 1
        S = 2 * f - cfg["ntaps"]
        n_in = cfg["n"] // S * S
        return n_in, (n_in + cfg["deci"] - 1) // cfg["deci"] if op == "fftfilt_decim" else n_in
    if op == "fir":
        return cfg["n"], (cfg["n"] - cfg["ntaps"] + 1) // cfg["deci"]
    if op == "fir_demod":
        return cfg["n"] * cfg["nchan"], ((cfg["n"] - cfg["ntaps"] + 1) // cfg["deci"] - 1) * cfg["nchan"]
    if op == "resample":
        return cfg["n"], -(-(cfg["n"] * cfg["interp"]) // cfg["deci"])
    return cfg["n"], cfg["n"]


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; rustradio
    cannot be compiled here), on every host core: one rustradio block processes one stream on one thread
    (FftFilter::work has no threading, src/fft_filter.rs:291), so the all-cores arm runs one independent
    stream per core, the way MTGraph would run as many independent chains."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]()
    threads = os.cpu_count() or 1
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    res = cpu_baseline(cfg, threads=threads, budget_s=max(5.0, min(60.0, 3.0 * steps)))
    n_in, n_out = static_sizes(cfg)
    W = Workload(units=n_in, n_in=n_in, n_out=n_out, parallelism=f"independent stream per GPU x{max(1, args.gpus)}")
    line = {
        "impl": "reference",
        "metric": "Msamples/s (c32) FIR/FftFilter/resampler at 1/2/4/8 B200; % of roofline",
        "value": res["value"], "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if cfg["dtype"] == "f32" else "c32 (complex f32)", "data": "synthetic",
        "config": config_block(cfg, W, max(1, args.gpus)),
        "cpu_baseline": res,
        "e2e": {"value": res["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--shard", default="capture", choices=["capture", "time", "channel"],
                    help="N>1 headline: independent capture per GPU (weak, default), one capture split by time segment with the halo "
                         "read over NVLink peer memory (strong; c2, c4), or config 3's channels split across the GPUs (strong)")
    ap.add_argument("--n", type=int, default=0, help="override the config's sample count (experiments only; the line's workload string says so)")
    ap.add_argument("--headline-only", action="store_true", help="skip the configs / splits sub-records")
    ap.add_argument("--no-ceiling", action="store_true", help="skip the bare-copy ceiling of the e2e leg")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of back-to-back headline steps for the `sustained` sub-record (0 = off)")
    ap.add_argument("--sub-cpu-budget", type=float, default=4.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-budget", type=float, default=10.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
