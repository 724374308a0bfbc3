"""Multi-GPU sharding of the filtering path: pure integer logic, no compute.

Two ways the path shards, neither needs a data-path collective:
  * by channel / capture: rank r owns channels [r*C/W, (r+1)*C/W);
  * by time segment: split the OUTPUT range evenly; a shard needs its inputs plus a left halo of
    ntaps-1 samples (FIR / FftFilter), 1 sample (demod) or none (resampler).
"""
from __future__ import annotations

from dataclasses import dataclass


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [lo, hi) share of `total` units for `rank` (first total % world ranks get one extra)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@dataclass(frozen=True)
class Segment:
    out_lo: int      # first output index owned
    out_hi: int      # one past the last output index owned
    in_lo: int       # first input sample needed (includes the halo)
    in_hi: int       # one past the last input sample needed


def fir_segment(n_in: int, ntaps: int, deci: int, world: int, rank: int) -> Segment:
    """FirFilter (src/fir.rs:192-197): output i reads inputs [i*deci, i*deci + ntaps); a shard of m outputs
    is handed the reference's `need` = m*deci + ntaps - 1 samples (src/fir.rs:502-507), which always exist."""
    n_out = 0 if n_in < ntaps + deci - 1 else (n_in - ntaps + 1) // deci
    lo, hi = shard_range(n_out, world, rank)
    if hi == lo:
        return Segment(lo, hi, 0, 0)
    return Segment(lo, hi, lo * deci, hi * deci + ntaps - 1)


def fftfilt_segment(n_in: int, ntaps: int, world: int, rank: int) -> Segment:
    """FftFilter (full convolution from n = 0, whole reference blocks only): output n reads inputs
    [n - ntaps + 1, n]; the first shard's halo is the zero initial state (src/fft_filter.rs:270)."""
    f = 1
    while f < ntaps:
        f <<= 1
    s = 2 * f - ntaps                      # nsamples, src/fft_filter.rs:36-42,262-263
    n_out = (n_in // s) * s
    lo, hi = shard_range(n_out, world, rank)
    if hi == lo:
        return Segment(lo, hi, 0, 0)
    return Segment(lo, hi, max(0, lo - (ntaps - 1)), hi)


def resampler_segment(n_in: int, interp: int, deci: int, world: int, rank: int) -> Segment:
    """RationalResampler (src/rational_resampler.rs:181-198): out[k] = in[floor(k*deci/interp)], no halo."""
    from math import gcd
    g = gcd(interp, deci)
    interp, deci = interp // g, deci // g
    n_out = -(-(n_in * interp) // deci)
    lo, hi = shard_range(n_out, world, rank)
    if hi == lo:
        return Segment(lo, hi, 0, 0)
    return Segment(lo, hi, (lo * deci) // interp, ((hi - 1) * deci) // interp + 1)


def demod_segment(n_in: int, world: int, rank: int) -> Segment:
    """QuadratureDemod (src/quadrature_demod.rs:71-73): out[t] reads in[t], in[t+1]."""
    lo, hi = shard_range(max(n_in - 1, 0), world, rank)
    if hi == lo:
        return Segment(lo, hi, 0, 0)
    return Segment(lo, hi, lo, hi + 1)


def resampler_shard_counter(seg: Segment, interp: int, deci: int) -> int:
    """The reference's `counter` state (src/rational_resampler.rs:101-105) at which a shard's work() loop
    must start so that its first output is global output seg.out_lo: the loop adds `interp` per input and
    emits while counter > 0, so after in_lo inputs and out_lo outputs counter = in_lo*interp - out_lo*deci,
    always in (-interp, 0] (a value <= -deci just means the first input sample's earlier outputs belong to
    the previous shard)."""
    from math import gcd
    g = gcd(interp, deci)
    return seg.in_lo * (interp // g) - seg.out_lo * (deci // g)
