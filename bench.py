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
    return dict(name="c1", op="fir", ntaps=64, deci=1, n=1 << 24, cutoff=0.1, dtype="c32",
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
    return dict(name="c3", op="fir_demod", ntaps=255, deci=10, nchan=1024, n=240_000, dtype="c32",
                desc="rtl_fm channelizer: 1024 ch x 240000 c32 (0.1 s @2.4 Msps), 255-tap /10 FIR + QuadratureDemod fused")


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
def run_gpu(args):
    import torch
    import rustradio_b200 as R

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = local
    cfg = CONFIGS[args.config]()
    stream = torch.cuda.current_stream().cuda_stream
    seed = SEED + (int(cfg["name"][1]) if cfg["name"][1].isdigit() else 9) + 1000 * rank      # every rank filters its own capture
    u8 = bool(cfg.get("in_u8"))

    def synth_input(nsamp):
        """Device-resident synthetic input: c32 white noise, or (u8 ingest) the same noise quantised
        to RTL-SDR bytes by an untimed setup step."""
        if not u8:
            t = torch.empty(2 * nsamp, dtype=torch.float32, device=f"cuda:{dev}")
            R.synth_f32(t, seed, 0, 2 * nsamp, dev, stream)
            return t
        out = torch.empty(2 * nsamp, dtype=torch.uint8, device=f"cuda:{dev}")
        chunk = 1 << 27
        tmp = torch.empty(min(chunk, 2 * nsamp), dtype=torch.float32, device=f"cuda:{dev}")
        for o in range(0, 2 * nsamp, chunk):
            m = min(chunk, 2 * nsamp - o)
            R.synth_f32(tmp, seed, o, m, dev, stream)
            torch.cuda.synchronize()
            out[o:o + m] = ((tmp[:m] + 1.0) * 128.0).floor_().clamp_(0, 255).to(torch.uint8)
        return out

    # ---- build the op and its device-resident input ----
    op = cfg["op"]
    scaling = "weak"
    if op == "fftfilt" and args.shard == "time" and world > 1:
        # ONE 2^28-sample capture split by time segment across the ranks (strong scaling): rank r
        # filters outputs [lo, hi) and receives its left halo (ntaps-1 samples) from rank r-1 over
        # NVLink (NCCL send/recv) every step; the halo becomes the filter's carried history.
        from rustradio_b200 import shard as S
        n = cfg["n"]
        f = R.FftFilt(taps_for(cfg), device=dev)
        seg = S.fftfilt_segment(n, cfg["ntaps"], world, rank)
        n_in = n_out = seg.out_hi - seg.out_lo
        T1 = cfg["ntaps"] - 1
        din = torch.empty(2 * n_in, dtype=torch.float32, device=f"cuda:{dev}")
        dout = torch.empty(2 * n_out, dtype=torch.float32, device=f"cuda:{dev}")
        halo = torch.zeros(2 * T1, dtype=torch.float32, device=f"cuda:{dev}")
        seed = SEED + int(cfg["name"][1])                  # one capture: same seed on every rank
        R.synth_f32(din, seed, 2 * seg.out_lo, 2 * n_in, dev, stream)
        tail = din[2 * (n_in - T1):]

        def step():
            ops = []
            if rank + 1 < world:
                ops.append(dist.P2POp(dist.isend, tail, rank + 1))
            if rank > 0:
                ops.append(dist.P2POp(dist.irecv, halo, rank - 1))
            for w in (dist.batch_isend_irecv(ops) if ops else []):
                w.wait()
            f.set_history(halo, T1, stream)                # rank 0: zeros = the stream's initial state
            f.run(din, n_in, dout, stream)
        launches_per_step = 2
        units = n_in
        scaling = "strong"
    elif op == "fftfilt":
        n = cfg["n"]
        f = R.FftFilt(taps_for(cfg), device=dev)
        n_in = (n // f.nsamples) * f.nsamples            # reference count rule (whole blocks)
        n_out = n_in
        din = torch.empty(2 * n, dtype=torch.float32, device=f"cuda:{dev}")
        dout = torch.empty(2 * n_out, dtype=torch.float32, device=f"cuda:{dev}")
        R.synth_f32(din, seed, 0, 2 * n, dev, stream)

        def step():
            f.run(din, n_in, dout, stream)
        launches_per_step = 2
        units = n_in
    elif op == "fftfilt_real":
        n = cfg["n"]
        f = R.FftFilt(taps_for(cfg).real.astype(np.float32), device=dev, real=True)
        n_in = n_out = (n // f.nsamples) * f.nsamples
        din = torch.empty(n, dtype=torch.float32, device=f"cuda:{dev}")
        dout = torch.empty(n_out, dtype=torch.float32, device=f"cuda:{dev}")
        R.synth_f32(din, seed, 0, n, dev, stream)

        def step():
            f.run(din, n_in, dout, stream)
        launches_per_step = 2
        units = n_in
    elif op == "fftfilt_decim":
        n = cfg["n"]
        f = R.FftFilt(taps_for(cfg), device=dev)
        n_in = (n // f.nsamples) * f.nsamples
        n_out = (n_in + cfg["deci"] - 1) // cfg["deci"]
        if u8:
            f.set_input_u8iq(True)
        din = synth_input(n)
        dout = torch.empty(2 * n_out, dtype=torch.float32, device=f"cuda:{dev}")

        def step():
            assert f.decim_run(din, n_in, cfg["deci"], 0, dout, stream) == n_out
        launches_per_step = 3
        units = n_in
    elif op == "fir":
        n = cfg["n"]
        f = R.Fir(taps_for(cfg), deci=cfg["deci"], device=dev)
        n_out = f.out_count(n)
        n_in = n
        fl = 1 if cfg["dtype"] == "f32" else 2                # floats per sample
        din = torch.empty(fl * n, dtype=torch.float32, device=f"cuda:{dev}")
        dout = torch.empty(fl * n_out, dtype=torch.float32, device=f"cuda:{dev}")
        R.synth_f32(din, seed, 0, fl * n, dev, stream)

        def step():
            f.run(din, n, dout, n_out, stream)
        launches_per_step = 1
        units = n_in
    elif op == "fir_demod":
        n, nchan = cfg["n"], cfg["nchan"] // (1 if world == 1 else 1)
        f = R.Fir(taps_for(cfg), deci=cfg["deci"], device=dev)
        out_n = f.out_count(n)
        need = (out_n - 1) * cfg["deci"] + cfg["ntaps"]
        if u8:
            f.set_input_u8iq(True)
        din = synth_input(n * nchan)
        dout = torch.empty((out_n - 1) * nchan, dtype=torch.float32, device=f"cuda:{dev}")

        def step():
            f.demod_run_batch(din, n, need, 1.0, dout, out_n - 1, out_n, nchan, stream)
        launches_per_step = 1
        n_in, n_out = n * nchan, (out_n - 1) * nchan
        units = n_in
    elif op == "decode":
        n = n_in = n_out = cfg["n"]
        u8 = True
        din = synth_input(n)
        dout = torch.empty(2 * n, dtype=torch.float32, device=f"cuda:{dev}")

        def step():
            R.rtlsdr_decode(din, 2 * n, dout, dev, stream)
        launches_per_step = 1
        units = n
    elif op == "fft":
        n = n_in = n_out = cfg["n"]
        f = R.Fft(cfg["size"], device=dev)
        din = synth_input(n)
        dout = torch.empty(2 * n, dtype=torch.float32, device=f"cuda:{dev}")

        def step():
            f.run(din, n // cfg["size"], dout, stream)
        launches_per_step = 1
        units = n
    elif op in ("hilbert", "mulconst", "mag2", "tee", "iqbalance"):
        n = n_in = n_out = units = cfg["n"]
        fin = 1 if op == "hilbert" else 2                       # floats per input sample
        fout = 1 if op == "mag2" else 2
        din = torch.empty(fin * n, dtype=torch.float32, device=f"cuda:{dev}")
        dout = torch.empty(fout * n, dtype=torch.float32, device=f"cuda:{dev}")
        R.synth_f32(din, seed, 0, fin * n, dev, stream)
        L = R.lib()
        launches_per_step = 1
        if op == "hilbert":
            f = R.Hilbert(cfg["ntaps"], device=dev)
            launches_per_step = 2

            def step():
                f.run(din, n, dout, stream)
        elif op == "iqbalance":
            f = R.IqBalance(R.iq_balance_alpha_from_tau(2_400_000, 0.2), device=dev)
            launches_per_step = 3

            def step():
                f.run(din, n, dout, stream)
        elif op == "mulconst":
            def step():
                assert L.rrc_multiply_const_c32_run(dev, din.data_ptr(), n, 0.3, -1.7, dout.data_ptr(), stream) == 0
        elif op == "mag2":
            def step():
                assert L.rrc_complex_to_mag2_run(dev, din.data_ptr(), n, dout.data_ptr(), stream) == 0
        else:
            dout2 = torch.empty(2 * n, dtype=torch.float32, device=f"cuda:{dev}")

            def step():
                assert L.rrc_tee_run(dev, din.data_ptr(), 8 * n, dout.data_ptr(), dout2.data_ptr(), stream) == 0
    elif op == "resample":
        n = cfg["n"]
        f = R.Resampler(4, cfg["interp"], cfg["deci"], device=dev)
        n_out = (n * cfg["interp"] + cfg["deci"] - 1) // cfg["deci"]
        n_in = n
        din = torch.empty(n, dtype=torch.float32, device=f"cuda:{dev}")
        dout = torch.empty(n_out + 16, dtype=torch.float32, device=f"cuda:{dev}")
        R.synth_f32(din, seed, 0, n, dev, stream)

        def step():
            f.reset()
            c, p, w = f.run(din, n, dout, n_out + 16, stream)
            assert (c, p) == (n, n_out)
        launches_per_step = 1
        units = n_in
    else:
        raise SystemExit(f"unknown op {op}")

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    l0 = R.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = R.launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{dev}")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    if scaling == "strong":
        tot = torch.tensor([float(units)], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(tot)
        value = float(tot.item()) / (ms_per_step * 1e-3) / 1e6
    else:
        value = units * world / (ms_per_step * 1e-3) / 1e6      # Msamples/s, whole job

    # ---- end to end: host buffers through *_run_host ----
    e2e = None
    if not args.no_e2e and op in ("fftfilt", "fir", "fftfilt_decim", "fft", "fftfilt_real") and scaling == "weak":
        real = op == "fftfilt_real" or (op == "fir" and cfg["dtype"] == "f32")
        ib = 2 if u8 else 4 if real else 8
        ob = 4 if real else 8
        hin = R.PinnedBuffer(np.uint8 if u8 else np.float32 if real else np.complex64, cfg["n"] * (2 if u8 else 1))
        hout = R.PinnedBuffer(np.float32 if real else np.complex64, n_out)
        R.lib().rrc_memcpy_d2h(dev, hin.ptr, din.data_ptr(), cfg["n"] * ib, stream)
        torch.cuda.synchronize()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        run_host = (lambda: f.decim_run_host(hin, cfg["deci"], hout)) if op == "fftfilt_decim" else (lambda: f.run_host(hin, hout))
        for _ in range(2):
            f.reset() if op in ("fftfilt", "fftfilt_decim", "fftfilt_real") else None
            run_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            if op in ("fftfilt", "fftfilt_decim", "fftfilt_real"):
                f.reset()
            got = run_host()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{dev}")
        if dist:
            dist.barrier()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {"value": units * world / (dt / e2e_steps) / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(ib * (n_in if op not in ("fir",) else cfg["n"])), "d2h_bytes_per_step": int(ob * len(got)),
               "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3, "timer": "host wall clock around rrc_*_run_host (returns after D2H completes)"}
        hin.free(); hout.free()

    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = json.loads(peaks_path.read_text())["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ab = alg_bytes(cfg, n_in, n_out)
    achieved = ab / (ms_per_step * 1e-3) / 1e9
    traffic_path = ROOT / "profiles" / "traffic.json"
    traffic = None
    if traffic_path.exists():
        traffic = json.loads(traffic_path.read_text()).get(cfg["name"])
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": ab,
                "kernel": {"fftfilt": "fftfilt_tma_kernel (TMA-staged input; RRC_FFTFILT_VARIANT=32 selects the LDG kernel)", "fir": "fir_poly_kernel<float2,float,1,false,16>", "fir_demod": "fir_rt_kernel<10,DEMOD,8,2>",
                           "resample": "resample_kernel", "decode": "rtlsdr_decode_kernel", "fft": "fftstream_kernel<10>", "fftfilt_real": "fftfilt_kernel (real-stream mode)",
                           "fftfilt_decim": "fftfilt_fold_kernel<4> (65536-point, 4-CTA cluster) + history update",
                           "hilbert": "hilbert_half_kernel + history update", "mulconst": "map_kernel<MAP_MUL_C32>", "mag2": "mag2_kernel",
                           "tee": "tee_kernel<uint4>", "iqbalance": "iq_tile_kernel<false> + iq_carry_kernel + iq_tile_kernel<true> (input read twice: 24 B/sample of traffic vs 16 algorithmic)"}[op],
                "duration_ms": ms_per_step,
                "note": "duration = CUDA-event time of the whole step on the launching stream / steps; the step is this one kernel"
                        + (" plus a <3 us history-update kernel" if op == "fftfilt" else "")}

    if op in ("fir", "fir_demod") and f.uses_tensor_cores:
        # Declared: the real-tap c32 FIR runs as a block-scaled fp16x3 Toeplitz product on the tensor cores (fir_tc.cuh).
        walk = cfg["deci"] in (1, 2, 4) and 7 * cfg["deci"] + cfg["ntaps"] <= 320
        roofline["kernel"] = (("fir_tcf_kernel<KS,D>" if cfg["dtype"] == "f32" else "fir_tc1_kernel<KS,DEMOD,U8,D>") if walk else "fir_tc_kernel") + \
            " (block-scaled fp16x3 Toeplitz product on the tensor cores, mma.m16n8k16 + ldmatrix" + (", fused demod epilogue)" if op == "fir_demod" else ")")
        ks = (7 * cfg["deci"] + cfg["ntaps"] + 15) // 16             # k-steps of 16 at 8 outputs per block-row (lower bound)
        nout_fir = n_out + (cfg.get("nchan", 0) if op == "fir_demod" else 0)
        mmas = 3 * ks * nout_fir / 64                                 # three m16n8k16 per k-step per 64 complex outputs
        mma_peak = 148 * 0.46 * 1.965e9                               # measured, profiles/r01_microbench_hmma_rate.txt
        roofline["tensor"] = {"mma_m16n8k16_per_launch": mmas, "achieved_mma_per_s": mmas / (ms_per_step * 1e-3),
                              "peak_mma_per_s": mma_peak, "frac": mmas / (ms_per_step * 1e-3) / mma_peak,
                              "peak_source": "measured mma.sync m16n8k16 issue rate, 0.46 per clk per SM (tools/microbench/hmma_rate.cu)"}
    # FP32 side of the roofline: algorithmic flops of the reference formulation against the
    # FP32 FMA rate MEASURED on this pool's B200 (tools/microbench/fp32_pipes.cu: 125 lanes/clk/SM).
    flops = alg_flops(cfg, n_in, n_out, f if op in ("fir", "fir_demod") else None)
    if flops:
        fp_peak = 148 * 125.0 * 2 * 1.965e9 / 1e12
        roofline["fp32"] = {"algorithmic_flops_per_launch": flops, "achieved_tflops": flops / (ms_per_step * 1e-3) / 1e12,
                            "peak_tflops": fp_peak, "frac": flops / (ms_per_step * 1e-3) / 1e12 / fp_peak,
                            "peak_source": "measured FFMA issue rate (profiles/r01_microbench_fp32_pipes.txt) x 2 flop"}

    cpu = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_baseline(cfg, threads=1, budget_s=args.cpu_budget)

    line = {
        "metric": "Msamples/s (c32) FIR/FftFilter/resampler at 1/2/4/8 B200; % of roofline",
        "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32" if cfg["dtype"] == "f32" else "c32 (complex f32)", "data": "synthetic",
        "config": {"workload": cfg["desc"], "name": cfg["name"], "samples_per_gpu_per_step": int(units),
                   "outputs_per_gpu_per_step": int(n_out), "parallelism": (f"independent stream per GPU x{world}" if scaling == "weak" else
                                   f"one capture, time-segment sharded x{world}, (ntaps-1)-sample halo by NCCL P2P"),
                   "l2_policy": "inputs larger than L2 (>= 0.5 GiB per step vs 126 MB L2)" if ab > 4e8 else "input 128 MiB ~ L2 size; see DESIGN.md",
                   "timer": "torch.cuda.Event on the launching stream, max over ranks"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


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


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; rustradio
    cannot be compiled here).  One rustradio block processes one stream on one thread
    (FftFilter::work has no threading, src/fft_filter.rs:291), so threads = streams = --gpus."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]()
    streams = max(1, args.gpus)
    threads = min(streams, os.cpu_count() or 1)
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    res = cpu_baseline(cfg, threads=threads, budget_s=max(5.0, min(60.0, 3.0 * steps)))
    line = {
        "impl": "reference",
        "metric": "Msamples/s (c32) FIR/FftFilter/resampler at 1/2/4/8 B200; % of roofline",
        "value": res["value"], "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if cfg["dtype"] == "f32" else "c32 (complex f32)", "data": "synthetic",
        "config": {"workload": cfg["desc"], "name": cfg["name"], "parallelism": f"{threads} host thread(s), one stream each"},
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
    ap.add_argument("--shard", default="capture", choices=["capture", "time"],
                    help="N>1: independent capture per GPU (weak, default) or one capture split by time segment with halo exchange (strong)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
