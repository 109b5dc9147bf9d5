"""vs_flow_analyze_batch (F0 / local jitter / local shimmer of generated glottal flow) against the numpy
restatement of its definitions (tests/analysis_ref.py) and against what the generator itself logged per pitch period."""
import math

import numpy as np
import pytest

import analysis_ref as ar

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vs():
    from voice_synth_b200 import api
    return api


@pytest.fixture(scope="module")
def ctx(vs):
    c = vs.Context()
    yield c
    c.close()


def _voices(vs, n, seed, noise=False):
    rng = np.random.default_rng(seed)
    args, seeds = [], []
    for i in range(n):
        f0 = float(rng.uniform(70, 320))
        a = ["-d", f"{rng.uniform(0.5, 1.5):.3f}", "-f", f"{f0:.2f}", "-g", f"{f0 + 30:.2f}"]
        if i % 4 != 0:
            a += ["-j", f"{rng.uniform(0.2, 4):.2f}"]
        if i % 4 != 1:
            a += ["-s", f"{rng.uniform(0.5, 12):.2f}"]
        if noise:
            a += ["-n", f"{rng.uniform(30, 45):.1f}"]
        if rng.random() < 0.5:
            a += ["-a", str(int(rng.integers(3000, 18000)))]
        args.append(a)
        seeds.append(int(rng.integers(0, 2**32)))
    return vs.FlowParams.from_cli(args, seeds)


def _same(got, want, i):
    assert int(got["onsets"]) == want["onsets"] and int(got["cycles"]) == want["cycles"] and int(got["flags"]) == want["flags"], (i, got, want)
    for k in ("f0_hz", "jitter_pct", "shimmer_pct", "mean_period", "mean_peak"):
        assert math.isclose(float(got[k]), float(np.float32(want[k])), rel_tol=2e-6, abs_tol=1e-30), (i, k, got[k], want[k])


def test_analysis_matches_restatement_and_period_log(ctx, vs):
    """noise-free flow, thresholds at the closed-phase level: the GPU statistics equal the numpy restatement, the cycle
    lengths are the generator's own pitch periods and the peaks its amplitudes"""
    import torch
    n = 96
    p = _voices(vs, n, 11)
    flow, offs, ns, log = ctx.flowgen_batch(p, want_log=True)
    st = ctx.flow_analyze_batch(flow, ns, offsets=offs, fs=p.fs)
    dev = torch.from_numpy(flow).cuda()
    st_dev = ctx.flow_analyze_batch(dev, ns, offsets=offs, fs=p.fs)
    assert st.tobytes() == st_dev.tobytes()
    for i in range(n):
        x = flow[int(offs[i]): int(offs[i]) + int(ns[i])]
        want = ar.stats(x, fs=int(p.fs[i]))
        _same(st[i], want, i)
        # against the generator's log: an onset is the second sample of a pitch period (x[start] = ceil(A*h[0]) = 0)
        o = ar.onsets(x)
        T_log = log[i]["T"].astype(np.int64)
        nc = len(o) - 1
        assert nc >= len(T_log) - 2 and np.array_equal(np.diff(o), T_log[:nc]), i
        peaks = np.array([x[o[k]: o[k + 1]].max() for k in range(nc)])
        assert np.array_equal(peaks, np.ceil(log[i]["A"][:nc].astype(np.float64)).astype(np.int64)), i
        # and against what was asked for
        has_j, has_s = bool(p.flags[i] & vs.VS_F_JITTER) and p.jitter[i] > 0, bool(p.flags[i] & vs.VS_F_SHIMMER) and p.shimmer[i] > 0
        P = int(np.float32(int(p.fs[i])) / np.float32(p.F0[i]))
        if not has_j:
            assert st[i]["jitter_pct"] == 0.0 and st[i]["mean_period"] == P
        else:
            assert st[i]["jitter_pct"] > 0.0 and 0.8 * P <= st[i]["mean_period"] <= 1.2 * P
        assert (st[i]["shimmer_pct"] > 0.0) == has_s


def test_analysis_noisy_flow_ragged_rows(ctx, vs):
    """glottal noise (30-45 dB): thresholds at 8 % / 15 % of the nominal amplitude -- above the noise, below the weakest
    shimmered pulse (0.2 x amplitude); rows at odd offsets with gaps"""
    n = 48
    p = _voices(vs, n, 12, noise=True)
    ns = vs.flow_nsamples(p)
    rng = np.random.default_rng(5)
    gaps = rng.integers(0, 50, n).astype(np.uint64)
    offs = (np.concatenate([[0], np.cumsum(ns + gaps)[:-1]]) + gaps).astype(np.uint64)
    flow = np.zeros(int(offs[-1] + ns[-1]), dtype=np.int16)
    _, _, _, log = ctx.flowgen_batch(p, out=flow, offsets=offs, want_log=True)
    lo, hi = (p.amp * 0.08).astype(np.int16), (p.amp * 0.15).astype(np.int16)
    st = ctx.flow_analyze_batch(flow, ns, offsets=offs, fs=p.fs, lo=lo, hi=hi)
    for i in range(n):
        x = flow[int(offs[i]): int(offs[i]) + int(ns[i])]
        _same(st[i], ar.stats(x, fs=int(p.fs[i]), lo=int(lo[i]), hi=int(hi[i])), i)
        assert abs(int(st[i]["cycles"]) - len(log[i])) <= 2, i          # one trigger per pitch period
        assert abs(float(st[i]["mean_period"]) - float(log[i]["T"].mean())) < 0.02 * float(log[i]["T"].mean())


def test_analysis_overflow_and_empty(ctx, vs):
    """thresholds inside an oscillation: more onsets than the scratch holds -> flagged, no statistics; silence: no onsets"""
    n = 4096
    x = np.zeros(3 * n, dtype=np.int16)
    x[:n:2] = 5
    x[2 * n + 100] = 7
    st = ctx.flow_analyze_batch(x, [n, n, n])
    assert st[0]["flags"] == 1 and st[0]["cycles"] == 0 and st[0]["onsets"] == n // 2
    assert st[1]["flags"] == 0 and st[1]["onsets"] == 0 and st[1]["cycles"] == 0 and st[1]["f0_hz"] == 0.0
    assert st[2]["onsets"] == 1 and st[2]["cycles"] == 0
    with pytest.raises(vs.VsError):
        ctx.flow_analyze_batch(x, [n], lo=3, hi=2)
