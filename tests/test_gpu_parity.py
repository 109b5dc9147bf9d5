"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed
fixtures made from the unmodified reference binaries.  Run on the B200 box: pytest -m gpu."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def vs():
    from voice_synth_b200 import api
    return api


@pytest.fixture(scope="module")
def ctx(vs):
    c = vs.Context()
    yield c
    c.close()


def _oracle_par(oracle, vs, p, i):
    q = oracle.FlowPar()
    for k in ("dur", "jitter", "cq", "K", "F0", "DC", "noise", "Kvar", "shimmer"):
        setattr(q, k, float(getattr(p, k)[i]))
    q.Fg = 0.0
    q.fs, q.amp, q.seed = int(p.fs[i]), int(p.amp[i]), int(p.seed[i])
    q.has_jitter = int(bool(p.flags[i] & vs.VS_F_JITTER))
    q.has_shimmer = int(bool(p.flags[i] & vs.VS_F_SHIMMER))
    q.has_noise = int(bool(p.flags[i] & vs.VS_F_NOISE))
    return q


def _golden_flow_params(golden, vs):
    cases = golden["cases"]
    p = vs.FlowParams.from_cli([c["args"] for c in cases], [c["seed"] for c in cases])
    return cases, p


LOG_KEYS = ("T", "T2", "T3", "T4", "A", "Knew", "S", "ndraws", "ndw", "x_pow", "w_pow", "start")


def test_flowgen_matches_reference_fixtures(ctx, vs, golden, oracle):
    """bit-exact int16 flow for every knob the reference has, against hashes of the reference binaries' PCM"""
    cases, p = _golden_flow_params(golden, vs)
    out, offs, ns, logs = ctx.flowgen_batch(p, want_log=True)
    for i, c in enumerate(cases):
        pcm = out[int(offs[i]): int(offs[i]) + int(ns[i])]
        assert int(ns[i]) == c["n"]
        assert pcm[:16].tolist() == c["head"], c["name"]
        assert sha(pcm) == c["sha256"], (c["name"], c["seed"])
        # perturbation sequences, pulse boundaries, draw counts: bit-exact against the oracle's log
        par = _oracle_par(oracle, vs, p, i)
        _, olog = oracle.flowgen(par, want_log=True)
        assert len(logs[i]) == len(olog), c["name"]
        for k in LOG_KEYS:
            a, b = logs[i][k], olog[k]
            assert np.array_equal(a, b, equal_nan=True), (c["name"], k, np.flatnonzero(a != b)[:5])


def test_flowgen_chunked_equals_unchunked(ctx, vs, golden):
    cases, p = _golden_flow_params(golden, vs)
    ref, offs, ns = ctx.flowgen_batch(p)
    for L in (64, 1000, 4096):
        ctx.set_option(vs.OPT_CHUNK_SAMPLES, L)
        out, _, _ = ctx.flowgen_batch(p)
        assert np.array_equal(out, ref), L
    ctx.set_option(1, 0)


def test_flowgen_unaligned_rows(ctx, vs, golden):
    """rows at odd sample offsets (16-byte phase 1..7) with guard gaps that must stay untouched"""
    cases, p = _golden_flow_params(golden, vs)
    ns = vs.flow_nsamples(p)
    offs = np.zeros(p.n, dtype=np.uint64)
    pos = 3
    for i in range(p.n):
        offs[i] = pos
        pos += int(ns[i]) + 1 + (i % 7)
    buf = np.full(pos + 8, -12345, dtype=np.int16)
    ctx.set_option(1, 512)
    ctx.flowgen_batch(p, out=buf, offsets=offs)
    ctx.set_option(1, 0)
    mask = np.ones(buf.size, dtype=bool)
    for i, c in enumerate(cases):
        seg = buf[int(offs[i]): int(offs[i]) + int(ns[i])]
        mask[int(offs[i]): int(offs[i]) + int(ns[i])] = False
        assert sha(seg) == c["sha256"], c["name"]
    assert np.all(buf[mask] == -12345)


def test_vowel_filter_exact_mode_is_bit_exact(ctx, vs, golden, oracle):
    """unfused mul/sub in the reference's order: identical doubles, identical PCM, all 10 presets"""
    z = np.load(__import__("pathlib").Path(__file__).parent / "golden" / "cfg1_seed42.npz")
    flow = z["flow"]
    n = 10
    flows = np.tile(flow, n)
    f = vs.FilterParams(n, "aiu1234567")
    ctx.set_option(3, 1)
    out, offs, raw = ctx.vowel_filter_batch(flows, [flow.size] * n, f, want_raw=True)
    ctx.set_option(3, 0)
    A = next(c for c in golden["cases"] if c["name"] == "A_cfg1" and c["seed"] == 42)
    for i, v in enumerate("aiu1234567"):
        seg = out[i * flow.size:(i + 1) * flow.size]
        want = next(x for x in A["vowels"] if x["preset"] == v and x["extra"] == "")
        assert sha(seg) == want["sha256"], v
        o_pcm, o_raw = oracle.vowel(flow, v, want_raw=True)
        assert np.array_equal(raw[i * flow.size:(i + 1) * flow.size], o_raw), v
    assert np.array_equal(out[:flow.size], z["vowel_a"])


@pytest.mark.parametrize("chunk", [0, -1, 3000])
def test_vowel_filter_fast_mode_tolerances(ctx, vs, oracle, chunk):
    """FMA-contracted FP64 recurrence (and chunked carries): <= 1e-5 before quantisation, +-1 LSB after"""
    z = np.load(__import__("pathlib").Path(__file__).parent / "golden" / "cfg1_seed42.npz")
    flow = z["flow"]
    n = 12
    f = vs.FilterParams(n, "aiu1234567ai", gain=10.0, pre=1.0)
    f.gain[10], f.pre[10] = 20.0, 0.5
    f.gain[11], f.pre[11] = 3.5, 0.0
    ctx.set_option(1, chunk)
    out, offs, raw = ctx.vowel_filter_batch(np.tile(flow, n), [flow.size] * n, f, want_raw=True)
    t = ctx.timing()
    ctx.set_option(1, 0)
    if chunk > 0:
        assert t["chunks"] > n
    worst = 0.0
    for i in range(n):
        o_pcm, o_raw = oracle.vowel(flow, chr(f.preset[i]), gain=float(f.gain[i]), pre=float(f.pre[i]), want_raw=True)
        seg, rseg = out[i * flow.size:(i + 1) * flow.size], raw[i * flow.size:(i + 1) * flow.size]
        worst = max(worst, float(np.abs(rseg - o_raw).max()))
        assert np.abs(rseg - o_raw).max() <= 1e-5, (i, chunk)
        assert np.abs(seg.astype(np.int32) - o_pcm.astype(np.int32)).max() <= 1, (i, chunk)
    print("max raw err", worst)


def test_synth_fused_matches_pipeline(ctx, vs, golden, oracle):
    """fused flow+filter against oracle flowgen -> oracle vowel for every golden flow case x a preset"""
    cases, p = _golden_flow_params(golden, vs)
    presets = "".join("aiu1234567"[i % 10] for i in range(p.n))
    f = vs.FilterParams(p.n, presets)
    ctx.set_option(3, 1)                                   # exact mode first: bit-exact end to end
    out, offs, ns, raw = ctx.synth_batch(p, f, want_raw=True)
    ctx.set_option(3, 0)
    fast, _, _, fraw = ctx.synth_batch(p, f, want_raw=True)
    ctx.set_option(1, 2048)
    chk, _, _, craw = ctx.synth_batch(p, f, want_raw=True)
    tchunks = ctx.timing()["chunks"]
    ctx.set_option(1, 0)
    assert tchunks > p.n
    for i, c in enumerate(cases):
        par = _oracle_par(oracle, vs, p, i)
        flow = oracle.flowgen(par)
        o_pcm, o_raw = oracle.vowel(flow, presets[i], want_raw=True)
        sl = slice(int(offs[i]), int(offs[i]) + int(ns[i]))
        assert np.array_equal(raw[sl], o_raw), c["name"]
        assert np.array_equal(out[sl], o_pcm), c["name"]
        for got, graw in ((fast, fraw), (chk, craw)):
            assert np.abs(graw[sl] - o_raw).max() <= 1e-5, c["name"]
            assert np.abs(got[sl].astype(np.int32) - o_pcm.astype(np.int32)).max() <= 1, c["name"]


def test_device_resident_buffers(ctx, vs):
    """device pointers in, device pointers out (what bench.py's `value` leg uses)"""
    import torch
    from voice_synth_b200 import workloads
    p, f = workloads.cfg2(n=256)
    host, offs, ns = ctx.synth_batch(p, f)
    dev = torch.zeros(host.size, dtype=torch.int16, device="cuda")
    ctx.synth_batch(p, f, out=dev)
    ctx.sync()
    assert np.array_equal(dev.cpu().numpy(), host)
