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


@pytest.mark.parametrize("plan_warps", [0, 1])
def test_both_plan_kernels_match_reference_fixtures(ctx, vs, golden, plan_warps):
    """the period table comes from one thread per stream or one warp per stream (VS_OPT_PLAN_WARPS): both must
    reproduce the reference PCM, unchunked and with chunks that start from RNG snapshots"""
    cases, p = _golden_flow_params(golden, vs)
    ctx.set_option(vs.OPT_PLAN_WARPS, plan_warps)
    try:
        for L in (0, 1000):
            ctx.set_option(vs.OPT_CHUNK_SAMPLES, L)
            out, offs, ns = ctx.flowgen_batch(p)
            for i, c in enumerate(cases):
                pcm = out[int(offs[i]): int(offs[i]) + int(ns[i])]
                assert sha(pcm) == c["sha256"], (c["name"], c["seed"], L)
    finally:
        ctx.set_option(vs.OPT_CHUNK_SAMPLES, 0)
        ctx.set_option(vs.OPT_PLAN_WARPS, -1)


def test_plan_kernels_agree_on_random_batches(ctx, vs):
    """512 random voices of cfg2 and 512 of cfg3 (glottal noise): identical flow from both plan kernels"""
    from voice_synth_b200 import workloads
    for make in (workloads.cfg2, workloads.cfg3):
        p, _ = make(n=512)
        outs = []
        for w in (0, 1):
            ctx.set_option(vs.OPT_PLAN_WARPS, w)
            out, _, _ = ctx.flowgen_batch(p)
            outs.append(out.copy())
        ctx.set_option(vs.OPT_PLAN_WARPS, -1)
        assert np.array_equal(outs[0], outs[1]), make.__name__


def test_steep_closure_wraps_like_the_reference_or_is_refused(ctx, vs, oracle):
    """-k 3..20: A*(K*cos - K + 1) runs far below -32768 after the first value under DC; the reference has left the
    falling branch by then (flowgen_shimmer.c:329), so nothing of the wrapped values may show.  -k 20000: the
    16-bit cast wraps before the DC test, the reference's output is an accident -> VS_ERANGE"""
    args = [["-d", "0.5", "-f", "120", "-k", "3", "-z", "0.5", "-j", "1", "-s", "3"],
            ["-d", "0.5", "-f", "150", "-g", "160", "-k", "20", "-l", "0.1"],
            ["-d", "0.5", "-f", "90", "-k", "6", "-s", "10", "-n", "15"]]
    p = vs.FlowParams.from_cli(args, [5, 6, 7])
    out, offs, ns = ctx.flowgen_batch(p)
    for i in range(p.n):
        want = oracle.flowgen(_oracle_par(oracle, vs, p, i))
        got = out[int(offs[i]): int(offs[i]) + int(ns[i])]
        assert np.array_equal(got, want), i
    bad = vs.FlowParams.from_cli([["-d", "0.5", "-f", "120", "-k", "20000"]], [1])
    with pytest.raises(vs.VsError) as e:
        ctx.flowgen_batch(bad)
    assert e.value.code == vs.VS_ERANGE


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


def test_vowel_output_noise_n1(ctx, vs, golden, oracle):
    """SURVEY 8f N1, `vowel -n`: per-frame output noise on the GPU, bit-exact against the reference fixtures"""
    z = np.load(__import__("pathlib").Path(__file__).parent / "golden" / "cfg1_seed42.npz")
    flow = z["flow"]
    A = next(c for c in golden["cases"] if c["name"] == "A_cfg1" and c["seed"] == 42)
    runs = [v for v in A["vowels"] if "-n" in v["extra"]]
    assert runs
    for v in runs:
        ex = v["extra"].split()
        opt = {ex[i]: float(ex[i + 1]) for i in range(0, len(ex), 2)}
        f = vs.FilterParams(1, v["preset"], gain=opt.get("-g", 10.0), pre=opt.get("-p", 1.0))
        ctx.set_option(vs.OPT_EXACT_FILTER, 1)             # identical filtered PCM first, then the noise must match bit for bit
        pcm, _ = ctx.vowel_filter_batch(flow, [flow.size], f)
        ctx.set_option(vs.OPT_EXACT_FILTER, 0)
        ctx.vowel_noise_batch(pcm, [flow.size], opt["-n"], 42)
        assert pcm[:16].tolist() == v["head"]
        assert sha(pcm) == v["sha256"], v
    # batch form on device memory, several sampling rates / seeds / a stream that is left alone
    import torch
    n = 6
    ns = np.array([22050, 5000, 1100, 1, 30000, 22050], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(ns + 3)[:-1]]).astype(np.uint64)
    rng = np.random.default_rng(4)
    base = rng.integers(-20000, 20000, int(offs[-1] + ns[-1]) + 8).astype(np.int16)
    dev = torch.from_numpy(base.copy()).cuda()
    snr = [30.0, 12.5, 5.0, 20.0, 0.0, 40.0]
    fs = [22050, 44100, 11025, 22050, 22050, 16000]
    seeds = [1, 2, 3, 4, 5, 4000000000]
    ctx.vowel_noise_batch(dev, ns, snr, seeds, offsets=offs, fs=fs)
    got = dev.cpu().numpy()
    want = base.copy()
    for i in range(n):
        seg = base[int(offs[i]): int(offs[i]) + int(ns[i])]
        if snr[i] > 0:
            # the oracle's vowel_noise = filter + noise; feed it through an identity-free path: emulate with its noise stage only
            want[int(offs[i]): int(offs[i]) + int(ns[i])] = _oracle_noise_only(oracle, seg, snr[i], seeds[i], fs[i])
    assert np.array_equal(got, want)


def _oracle_noise_only(oracle, pcm, snr_db, seed, fs):
    """noise stage of the oracle on given PCM (python restatement of oracle/vs_oracle.c:vso_vowel_noise's second half)"""
    import ctypes as C
    import math
    L = oracle.lib()
    g = oracle.Rng()
    L.vso_srandom(C.byref(g), C.c_uint32(seed))
    f32 = np.float32
    snr = f32(math.pow(10.0, float(f32(snr_db) / f32(10))))
    frame = 50 * (int(fs * 0.001 / 2.0) * 2)
    y = pcm.copy()
    for base in range(0, y.size, frame):
        ni = min(frame, y.size - base)
        aux = f32(0)
        for i in range(ni):
            aux = f32(aux + f32(f32(y[base + i]) * f32(y[base + i])))
        power = f32(aux / f32(ni))
        width = f32(math.sqrt(float(f32(f32(12) * power) / snr)))
        for i in range(ni):
            nv = f32(L.vso_random(C.byref(g)) / 2147483647.0)
            a = f32(float(width) * (float(nv) - 0.5))
            y[base + i] = L.vso_round2int(float(y[base + i]) + float(a))
    return y


def test_overlapping_rows_are_refused(ctx, vs):
    """VS_EOVERLAP: output rows that intersect (they would be written by different CTAs at once), and a filter
    running in place on device memory (time-chunks re-read flow the previous chunk has already overwritten)"""
    import torch
    p = vs.FlowParams.from_cli(["-d 1 -f 120 -j 1 -s 3"] * 3, [1, 2, 3])
    f = vs.FilterParams(3, "aiu")
    out = np.zeros(3 * 22050, dtype=np.int16)
    for offs in ([0, 22049, 44100], [0, 0, 22050], [100, 22050, 22000]):
        with pytest.raises(vs.VsError) as e:
            ctx.synth_batch(p, f, out=out, offsets=np.array(offs, dtype=np.uint64))
        assert e.value.code == vs.VS_EOVERLAP
    # touching rows are fine
    pcm, _, _ = ctx.synth_batch(p, f, out=out, offsets=np.array([0, 22050, 44100], dtype=np.uint64))
    flow = torch.zeros(3 * 22050, dtype=torch.int16, device="cuda")
    ctx.flowgen_batch(p, out=flow)
    ctx.sync()
    with pytest.raises(vs.VsError) as e:
        ctx.vowel_filter_batch(flow, [22050] * 3, f, out=flow)
    assert e.value.code == vs.VS_EOVERLAP
    dst = torch.zeros_like(flow)
    ctx.vowel_filter_batch(flow, [22050] * 3, f, out=dst)
    ctx.sync()
    assert int((dst.cpu().numpy().astype(np.int32) - pcm.astype(np.int32)).__abs__().max()) <= 1
