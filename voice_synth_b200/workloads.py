"""Synthetic parameter batches of BASELINE.json's configs, as FlowParams/FilterParams.

All values go through the same float32 conversions the reference CLI applies, so every stream is
also a valid command line for the unmodified reference tools (see `cli_args`)."""
import math

import numpy as np

from .api import FilterParams, FlowParams, VS_F_JITTER, VS_F_NOISE, VS_F_SHIMMER

PRESETS = "aiu1234567"


def _finish(n, dur, F0, jit_pct, shm_pct, seed, snr_db=None, fs=22050):
    f32 = np.float32
    p = FlowParams(n)
    p.dur[...] = f32(dur)
    p.F0[...] = np.asarray(F0, dtype=np.float32)
    p.jitter[...] = (np.asarray(jit_pct, dtype=np.float64) / 100.0).astype(np.float32)         # atof()/100.0 -> float
    p.shimmer[...] = np.asarray(shm_pct, dtype=np.float32) / f32(100)                            # float / 100
    p.fs[...] = fs
    p.seed[...] = np.asarray(seed, dtype=np.uint32)
    flags = np.full(n, VS_F_JITTER | VS_F_SHIMMER, dtype=np.uint8)
    if snr_db is not None:
        snr = np.broadcast_to(np.asarray(snr_db, dtype=np.float32), (n,))
        on = snr >= 0
        noise = np.array([f32(math.pow(10.0, float(f32(s) / f32(10)))) if s >= 0 else f32(0) for s in snr], dtype=np.float32)
        p.noise[...] = noise
        p.DC[on] = f32(0.25)                                                                     # -n sets par.DC = .25
        flags[on] |= VS_F_NOISE
    p.flags[...] = flags
    # the decimal values a command line would carry (cli_args); the float32 fields above are what the
    # reference's atof()/100 conversions make of them
    p.meta = {"jit_pct": np.broadcast_to(np.asarray(jit_pct, dtype=np.float64), (n,)).copy(),
              "shm_pct": np.broadcast_to(np.asarray(shm_pct, dtype=np.float64), (n,)).copy(),
              "snr_db": None if snr_db is None else np.broadcast_to(np.asarray(snr_db, dtype=np.float64), (n,)).copy()}
    return p


def cfg1():
    """single voice: flowgen_shimmer -d 1 -f 120 -j 1 -s 3 | vowel -v a, seed 42"""
    return _finish(1, 1.0, [120.0], [1.0], [3.0], [42]), FilterParams(1, "a")


def cfg2(n=4096, dur=1.0):
    """4096 streams x 1 s covering every vowel preset (the bench workload)."""
    s = np.arange(n)
    p = _finish(n, dur, 80.0 + (s % 64) * 2.5, (s % 8) * 0.5, (s % 16) * 0.5, 1000 + s)
    f = FilterParams(n)
    f.preset[...] = [ord(PRESETS[i % 10]) for i in s]
    return p, f


def cfg3(n=65536, dur=2.0, first=0):
    """jitter x shimmer x F0 grid, 2 s, glottal noise at 20 dB. `first` offsets the grid (sharding)."""
    s = np.arange(first, first + n)
    jit = (s % 16) * 0.2                       # 16 values 0..3 %
    shm = ((s // 16) % 16) * 0.5               # 16 values 0..7.5 %
    F0 = 80.0 + ((s // 256) % 256) * 0.5       # 256 values 80..207.5 Hz
    p = _finish(n, dur, F0, jit, shm, s, snr_db=20.0)
    f = FilterParams(n)
    f.preset[...] = [ord(PRESETS[i % 10]) for i in s]
    return p, f


def cfg4(dur=600.0):
    """single 10-minute stream, vowel /i/."""
    return _finish(1, dur, [120.0], [1.0], [3.0], [42]), FilterParams(1, "i")


def _hash32(x):
    x = np.asarray(x, dtype=np.uint64)
    x = (x ^ (x >> np.uint64(16))) * np.uint64(0x45D9F3B) & np.uint64(0xFFFFFFFF)
    x = (x ^ (x >> np.uint64(16))) * np.uint64(0x45D9F3B) & np.uint64(0xFFFFFFFF)
    return (x ^ (x >> np.uint64(16))) & np.uint64(0xFFFFFFFF)


def cfg5(n=1 << 20, first=0, dur=1.0):
    """corpus sweep: parameters from a counter-based hash of the utterance id."""
    uid = np.arange(first, first + n, dtype=np.uint64)
    h = [_hash32(uid * np.uint64(5) + np.uint64(k)) for k in range(5)]
    u = [hk.astype(np.float64) / 4294967296.0 for hk in h]
    F0 = np.round((80.0 + u[0] * 170.0) * 2) / 2            # 80..250 Hz in 0.5 Hz steps
    jit = np.round(u[1] * 30) / 10                            # 0..3 % in 0.1 steps
    shm = np.round(u[2] * 100) / 10                           # 0..10 %
    snr_sel = (h[3] % np.uint64(5)).astype(np.int64)          # off, 10, 20, 30, 40 dB
    snr = np.where(snr_sel == 0, -1.0, snr_sel * 10.0)
    p = _finish(n, dur, F0, jit, shm, uid.astype(np.uint32), snr_db=snr)
    f = FilterParams(n)
    f.preset[...] = np.frombuffer(PRESETS.encode(), dtype=np.uint8)[(h[4] % np.uint64(10)).astype(np.int64)]
    return p, f


def cli_args(p, i):
    """the reference command line (without -o) that yields stream i of FlowParams p"""
    m = p.meta
    a = ["-d", repr(float(p.dur[i])), "-f", repr(float(p.F0[i])), "-g", repr(max(125.0, float(p.F0[i]) + 5.0))]
    if p.flags[i] & VS_F_JITTER:
        a += ["-j", repr(round(float(m["jit_pct"][i]), 6))]
    if p.flags[i] & VS_F_SHIMMER:
        a += ["-s", repr(round(float(m["shm_pct"][i]), 6))]
    if p.flags[i] & VS_F_NOISE:
        a += ["-n", repr(round(float(m["snr_db"][i]), 6))]
    if int(p.fs[i]) != 22050:
        a += ["-r", str(int(p.fs[i]))]
    return a
