"""BASELINE.json's configurations at (or near) full size on the GPU: spot-checked against the oracle on a
seeded sample of streams, plus size-independent properties (idempotence, shard invariance, chunking
invariance of the flow, exact stream lengths)."""
import numpy as np
import pytest

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


def _opar(oracle, vs, p, i):
    q = oracle.FlowPar()
    for k in ("dur", "jitter", "cq", "K", "F0", "DC", "noise", "Kvar", "shimmer"):
        setattr(q, k, float(getattr(p, k)[i]))
    q.fs, q.amp, q.seed = int(p.fs[i]), int(p.amp[i]), int(p.seed[i])
    q.has_jitter, q.has_shimmer, q.has_noise = [int(bool(p.flags[i] & b)) for b in (1, 2, 4)]
    return q


def _check_sample(oracle, vs, p, f, pcm, offs, ns, idx, flow=None, foffs=None):
    worst = 0
    for i in idx:
        oflow = oracle.flowgen(_opar(oracle, vs, p, i))
        if flow is not None:
            assert np.array_equal(flow[int(foffs[i]): int(foffs[i]) + int(ns[i])], oflow), f"flow of stream {i}"
        want = oracle.vowel(oflow, chr(f.preset[i]), gain=float(f.gain[i]), pre=float(f.pre[i]))
        got = pcm[int(offs[i]): int(offs[i]) + int(ns[i])]
        d = int(np.abs(got.astype(np.int32) - want.astype(np.int32)).max())
        worst = max(worst, d)
        assert d <= 1, f"stream {i}: {d} LSB"
    return worst


def _check_all(oracle, vs, p, f, pcm, offs, ns, flow=None, foffs=None, label=""):
    """EVERY stream of the batch against the oracle, on all host cores (the oracle is a ctypes library: the GIL is
    released inside it).  Flow bit-exact, PCM within 1 LSB; returns (streams, samples, samples off by one LSB)."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    def one(i):
        oflow = oracle.flowgen(_opar(oracle, vs, p, i))
        flow_ok = True if flow is None else bool(np.array_equal(flow[int(foffs[i]): int(foffs[i]) + int(ns[i])], oflow))
        want = oracle.vowel(oflow, chr(f.preset[i]), gain=float(f.gain[i]), pre=float(f.pre[i]))
        got = pcm[int(offs[i]): int(offs[i]) + int(ns[i])]
        d = np.abs(got.astype(np.int32) - want.astype(np.int32))
        return flow_ok, int(d.max()), int((d > 0).sum()), int(d.size)

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        res = list(ex.map(one, range(p.n), chunksize=16))
    bad_flow = [i for i, r in enumerate(res) if not r[0]]
    bad_pcm = [(i, r[1]) for i, r in enumerate(res) if r[1] > 1]
    assert not bad_flow, f"{label}: flow differs from the oracle in streams {bad_flow[:10]} ({len(bad_flow)} in all)"
    assert not bad_pcm, f"{label}: PCM more than 1 LSB off in streams {bad_pcm[:10]} ({len(bad_pcm)} in all)"
    ones, total = sum(r[2] for r in res), sum(r[3] for r in res)
    print(f"{label}: {p.n} streams, {total} samples checked against the oracle, {ones} samples differ by 1 LSB")
    return p.n, total, ones


def test_cfg2_full_batch(ctx, vs, oracle):
    """4096 streams x 1 s, every preset, fused kernel, auto chunking (the bench workload): every stream checked"""
    from voice_synth_b200 import workloads
    p, f = workloads.cfg2()
    pcm, offs, ns = ctx.synth_batch(p, f)
    t = ctx.timing()
    assert t["samples"] == 4096 * 22050 and t["chunks"] > 4096
    assert t["render_path"] & 1, "the bench workload must run on the branch-free generator"
    flow, foffs, _ = ctx.flowgen_batch(p)
    _check_all(oracle, vs, p, f, pcm, offs, ns, flow, foffs, label="cfg2")
    # idempotence: the same call again gives the same bytes
    pcm2, _, _ = ctx.synth_batch(p, f)
    assert np.array_equal(pcm, pcm2)
    # shard invariance: synthesising a sub-range alone gives the same rows (what a multi-GPU split does)
    sub = np.arange(1000, 1512)
    spcm, soffs, sns = ctx.synth_batch(p.select(sub), f.select(sub))
    for k, i in enumerate(sub[::37]):
        kk = int(k * 37)
        a = spcm[int(soffs[kk]): int(soffs[kk]) + int(sns[kk])]
        b = pcm[int(offs[i]): int(offs[i]) + int(ns[i])]
        assert np.abs(a.astype(np.int32) - b.astype(np.int32)).max() <= 1      # chunk boundaries differ -> +-1 LSB allowed
    # flow is bit-exact regardless of how the batch is cut
    sflow, sfo, _ = ctx.flowgen_batch(p.select(sub))
    for k, i in enumerate(sub[::37]):
        kk = int(k * 37)
        assert np.array_equal(sflow[int(sfo[kk]): int(sfo[kk]) + int(sns[kk])], flow[int(foffs[i]): int(foffs[i]) + int(ns[i])])


def test_cfg3_noise_grid_shard(ctx, vs, oracle):
    """one GPU's shard of the jitter x shimmer x F0 grid: 8192 streams x 2 s with glottal noise"""
    from voice_synth_b200 import workloads
    p, f = workloads.cfg3(n=8192, first=3 * 8192)
    pcm, offs, ns = ctx.synth_batch(p, f)
    assert int(ns.sum()) == 8192 * 44100
    assert ctx.timing()["render_path"] & 3 == 3, "cfg3 must run on the branch-free generator with in-lane noise"
    flow, foffs, _ = ctx.flowgen_batch(p)
    _check_all(oracle, vs, p, f, pcm, offs, ns, flow, foffs, label="cfg3 shard")


def test_cfg4_ten_minute_stream(ctx, vs, oracle):
    """single 600 s stream, vowel /i/: 13 230 000 samples through the time-chunked filter"""
    from voice_synth_b200 import workloads
    p, f = workloads.cfg4()
    flow, _, ns = ctx.flowgen_batch(p)
    assert int(ns[0]) == 13230000
    oflow = oracle.flowgen(_opar(oracle, vs, p, 0))
    assert np.array_equal(flow, oflow)
    pcm, _, _ = ctx.synth_batch(p, f)
    assert ctx.timing()["chunks"] > 100
    want = oracle.vowel(oflow, "i")
    d = np.abs(pcm.astype(np.int32) - want.astype(np.int32))
    assert int(d.max()) <= 1
    print("cfg4: samples differing by 1 LSB:", int((d > 0).sum()), "of", d.size)
    # the stand-alone filter on the reference-exact flow gives the same
    out, _ = ctx.vowel_filter_batch(oflow, [oflow.size], f)
    assert int(np.abs(out.astype(np.int32) - want.astype(np.int32)).max()) <= 1


def test_cfg5_corpus_slice(ctx, vs, oracle):
    """a slice of the 1M-utterance sweep: mixed noise / no-noise streams, hashed parameters"""
    from voice_synth_b200 import workloads
    p, f = workloads.cfg5(n=16384, first=500000)
    pcm, offs, ns = ctx.synth_batch(p, f)
    flow, foffs, _ = ctx.flowgen_batch(p)
    _check_all(oracle, vs, p, f, pcm, offs, ns, flow, foffs, label="cfg5 slice")


def test_every_preset_through_one_two_and_five_chunks(ctx, vs, oracle):
    """the ten presets x chunk classes grid: each vowel filtered as one row, as two and as five time-chunks (carry
    warm-up differs per preset), fused and filter-only, against the oracle's waveform before and after quantisation"""
    n = 10
    p = vs.FlowParams.from_cli(["-d 1 -f 120 -j 1 -s 3"] * n, list(range(100, 100 + n)))
    f = vs.FilterParams(n, "aiu1234567")
    want = []
    for i in range(n):
        oflow = oracle.flowgen(_opar(oracle, vs, p, i))
        want.append((oflow,) + tuple(oracle.vowel(oflow, chr(f.preset[i]), want_raw=True)))
    for chunks, L in ((1, -1), (2, 11032), (5, 4416)):
        ctx.set_option(vs.OPT_CHUNK_SAMPLES, L)
        try:
            pcm, offs, ns, raw = ctx.synth_batch(p, f, want_raw=True)
            assert ctx.timing()["chunks"] == n * chunks
            pcm2, offs2, ns2 = ctx.synth_batch(p, f)                       # without raw output: integer pre-emphasis path
            flow_all = np.concatenate([w[0] for w in want])
            fout, foffs, fraw = ctx.vowel_filter_batch(flow_all, ns, f, in_offsets=np.arange(n, dtype=np.uint64) * 22050, want_raw=True)
        finally:
            ctx.set_option(vs.OPT_CHUNK_SAMPLES, 0)
        for i in range(n):
            oflow, opcm, oraw = want[i]
            for name, got, graw, o in (("fused", pcm, raw, offs), ("fused/int", pcm2, None, offs2), ("filter", fout, fraw, foffs)):
                sl = slice(int(o[i]), int(o[i]) + 22050)
                d = int(np.abs(got[sl].astype(np.int32) - opcm.astype(np.int32)).max())
                assert d <= 1, f"preset {chr(f.preset[i])}, {chunks} chunk(s), {name}: {d} LSB"
                if graw is not None:
                    e = float(np.abs(graw[sl] - oraw).max())
                    assert e <= 1e-5, f"preset {chr(f.preset[i])}, {chunks} chunk(s), {name}: pre-quantisation error {e}"


def test_fast_paths_fuzz(ctx, vs, oracle):
    """random voices inside the domain of the branch-free generator (no -z, noise without -l, default gain and
    pre-emphasis so that both commute to the integer input), with and without glottal noise, chunked and not:
    every stream against the oracle"""
    rng = np.random.default_rng(77)
    for noisy in (False, True):
        n = 200
        args, seeds = [], []
        for i in range(n):
            f0 = float(rng.uniform(75, 300))
            a = ["-d", f"{rng.uniform(0.5, 1.2):.3f}", "-f", f"{f0:.2f}", "-g", f"{f0 + 20:.2f}"]
            if rng.random() < 0.85:
                a += ["-j", f"{rng.uniform(0, 4):.2f}"]
            if rng.random() < 0.85:
                a += ["-s", f"{rng.uniform(0, 12):.2f}"]
            if noisy and rng.random() < 0.7:
                a += ["-n", f"{rng.uniform(5, 45):.1f}"]
            if rng.random() < 0.4:
                a += ["-c", f"{rng.uniform(0.4, 0.9):.3f}"]
            if rng.random() < 0.4:
                a += ["-k", f"{rng.uniform(0.5, 2):.3f}"]
            if rng.random() < 0.4:
                a += ["-a", str(int(rng.integers(2000, 18000)))]
            args.append(a)
            seeds.append(int(rng.integers(0, 2**32)))
        p = vs.FlowParams.from_cli(args, seeds)
        f = vs.FilterParams(n)
        f.preset[...] = [ord("aiu1234567"[int(k)]) for k in rng.integers(0, 10, n)]
        f.gain[...] = rng.integers(1, 15, n).astype(np.float32)
        for chunk in (0, 2048):
            ctx.set_option(vs.OPT_CHUNK_SAMPLES, chunk)
            try:
                flow, foffs, ns = ctx.flowgen_batch(p)
                pcm, offs, _ = ctx.synth_batch(p, f)
                t = ctx.timing()
            finally:
                ctx.set_option(vs.OPT_CHUNK_SAMPLES, 0)
            assert t["render_path"] & 1 and bool(t["render_path"] & 2) == noisy and (t["render_path"] >> 2) & 3 == 0
            _check_all(oracle, vs, p, f, pcm, offs, ns, flow, foffs, label=f"fast fuzz noise={noisy} chunk={chunk}")
            # the general generator must agree bit for bit on the flow and within 1 LSB on the PCM
            ctx.set_option(vs.OPT_SIMPLE_GEN, 1)
            try:
                flow2, _, _ = ctx.flowgen_batch(p)
                pcm2, _, _ = ctx.synth_batch(p, f)
                assert not ctx.timing()["render_path"] & 1
            finally:
                ctx.set_option(vs.OPT_SIMPLE_GEN, 0)
            assert np.array_equal(flow, flow2)
            assert int(np.abs(pcm.astype(np.int32) - pcm2.astype(np.int32)).max()) <= 1


def test_flow_rows_kernel_layouts(ctx, vs, oracle):
    """vs_flowgen_batch without glottal noise runs on the warp-per-row kernel (lanes along the row): ragged rows at
    odd offsets with gaps, pitch periods from 30 to 560 samples, streams shorter than one 256-sample step, DC offsets,
    host and device buffers at every phase of a 128-byte line, chunked and not -- every stream bit for bit"""
    import torch
    rng = np.random.default_rng(4242)
    n = 160
    args, seeds = [], []
    for i in range(n):
        f0 = float(np.exp(rng.uniform(np.log(50.0), np.log(580.0))))
        a = ["-d", f"{rng.uniform(0.5, 1.7):.4f}", "-f", f"{f0:.2f}", "-g", f"{f0 * 1.2 + 10:.2f}"]
        if rng.random() < 0.7:
            a += ["-j", f"{rng.uniform(0, 5):.2f}"]
        if rng.random() < 0.7:
            a += ["-s", f"{rng.uniform(0, 15):.2f}"]
        if rng.random() < 0.5:
            a += ["-c", f"{rng.uniform(0.2, 1.0):.3f}"]
        if rng.random() < 0.5:
            a += ["-k", f"{rng.uniform(0.5, 3):.3f}"]
        if rng.random() < 0.3:
            a += ["-l", f"{rng.uniform(0, 0.3):.3f}"]
        if rng.random() < 0.4:
            a += ["-a", str(int(rng.integers(1, 18000)))]
        args.append(a)
        seeds.append(int(rng.integers(0, 2**32)))
    p = vs.FlowParams.from_cli(args, seeds)
    # the ABI takes any positive duration (the tool's usage() wants 0.5 s): rows shorter than a step, than a line
    for i in range(0, n, 3):
        p.dur[i] = np.float32([0.0005, 0.003, 0.011, 0.05, 0.3][(i // 3) % 5] * rng.uniform(0.8, 1.2))
    ns = vs.flow_nsamples(p)
    assert int(ns.min()) < 32 and int(ns.max()) > 30000
    # ragged layout: a gap of 0..70 samples in front of every row, so rows start at every phase of a line
    gaps = rng.integers(0, 71, n).astype(np.uint64)
    offs = (np.concatenate([[0], np.cumsum(ns + gaps)[:-1]]) + gaps).astype(np.uint64)
    total = int(offs[-1] + ns[-1]) + 64
    want = [oracle.flowgen(_opar(oracle, vs, p, i)) for i in range(n)]
    for chunk in (0, 512, 1000, -1):
        ctx.set_option(vs.OPT_CHUNK_SAMPLES, chunk)
        try:
            for shift in (0, 1, 37):
                host = np.full(total + shift, -12345, dtype=np.int16)
                ctx.flowgen_batch(p, out=host[shift:], offsets=offs)
                t = ctx.timing()
                assert t["render_path"] & 32, "flow without noise must take the warp-per-row kernel"
                dev = torch.full((total + shift,), -12345, dtype=torch.int16, device="cuda")
                torch.cuda.synchronize()                      # (the library runs on its own streams)
                ctx.flowgen_batch(p, out=dev[shift:], offsets=offs)
                ctx.sync()
                got_dev = dev.cpu().numpy()
                used = np.zeros(total + shift, dtype=bool)
                for i in range(n):
                    a0 = shift + int(offs[i])
                    used[a0: a0 + int(ns[i])] = True
                    assert np.array_equal(host[a0: a0 + int(ns[i])], want[i]), f"host, chunk {chunk}, shift {shift}, stream {i}"
                    assert np.array_equal(got_dev[a0: a0 + int(ns[i])], want[i]), f"device, chunk {chunk}, shift {shift}, stream {i}"
                # nothing outside the rows is touched (device buffers are written in place)
                assert np.all(got_dev[~used] == -12345), f"device, chunk {chunk}, shift {shift}: wrote outside the rows"
        finally:
            ctx.set_option(vs.OPT_CHUNK_SAMPLES, 0)
    # the lane-per-row kernel (forced by the option) gives the same bytes
    ctx.set_option(vs.OPT_SIMPLE_GEN, 1)
    try:
        host = np.zeros(total, dtype=np.int16)
        ctx.flowgen_batch(p, out=host, offsets=offs)
        assert not ctx.timing()["render_path"] & 32
    finally:
        ctx.set_option(vs.OPT_SIMPLE_GEN, 0)
    for i in range(n):
        assert np.array_equal(host[int(offs[i]): int(offs[i]) + int(ns[i])], want[i])


def test_edge_shapes(ctx, vs, oracle):
    """ragged batch: shortest legal stream, very high and very low F0, a stream shorter than one window"""
    args = ["-d 0.5 -f 50 -g 60", "-d 0.5 -f 1000 -g 1100 -j 1 -s 1", "-d 0.5 -r 100 -f 50 -g 60 -s 2",
            "-d 0.5 -r 8000 -f 399 -g 400 -j 3 -n 15", "-d 1.25 -f 120 -c 0 -s 3", "-d 0.75 -f 120 -a 0 -j 1"]
    p = vs.FlowParams.from_cli(args, [9, 8, 7, 6, 5, 4])
    f = vs.FilterParams(p.n, "aiu123")
    flow, foffs, ns = ctx.flowgen_batch(p)
    pcm, offs, _ = ctx.synth_batch(p, f)
    _check_sample(oracle, vs, p, f, pcm, offs, ns, range(p.n), flow, foffs)
    assert int(ns[2]) == 50                                   # 0.5 s at 100 Hz: less than one render window
    # n == 0 streams are rejected, not silently skipped
    bad = vs.FlowParams(1, dur=0.0)
    with pytest.raises(vs.VsError):
        ctx.flowgen_batch(bad)
    with pytest.raises(vs.VsError):
        ctx.synth_batch(vs.FlowParams(1), vs.FilterParams(1, "x"))


def test_fresh_context_first_call(vs, oracle):
    """first call on a new context: the cosine tables of ~50 distinct T2 are uploaded right before the plan
    kernel (a missing stream dependency there once produced garbage tables)"""
    from voice_synth_b200 import workloads
    c = vs.Context()
    try:
        p, f = workloads.cfg5(n=24576, first=7)
        pcm, offs, ns = c.synth_batch(p, f)
        rng = np.random.default_rng(11)
        _check_sample(oracle, vs, p, f, pcm, offs, ns, sorted(set(rng.integers(0, p.n, 16).tolist())))
    finally:
        c.close()


def test_calls_of_different_shapes_interleave(ctx, vs, oracle):
    """alternate batches of different shapes and modes: exercises the parity-double-buffered descriptors,
    the plan cache and the chunk-plan re-upload logic"""
    from voice_synth_b200 import workloads
    pa, fa = workloads.cfg2(n=700)
    pb, fb = workloads.cfg3(n=300, first=1234)
    ref_a = ctx.synth_batch(pa, fa)[0].copy()
    ref_b = ctx.synth_batch(pb, fb)[0].copy()
    for _ in range(3):
        assert np.array_equal(ctx.synth_batch(pa, fa)[0], ref_a)
        ctx.flowgen_batch(pb)
        assert np.array_equal(ctx.synth_batch(pb, fb)[0], ref_b)
        assert np.array_equal(ctx.synth_batch(pb, fb)[0], ref_b)
        ctx.flowgen_batch(pa)


def test_multi_device_context(vs, oracle):
    """one vs_ctx over two GPUs: contiguous stream ranges per device, results identical to one device"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from voice_synth_b200 import workloads
    p, f = workloads.cfg2(n=1536)
    one = vs.Context(devices=[0])
    two = vs.Context(devices=[0, 1])
    try:
        a, offs, ns = one.synth_batch(p, f)
        b, _, _ = two.synth_batch(p, f)
        assert np.abs(a.astype(np.int32) - b.astype(np.int32)).max() <= 1      # chunking differs per device share
        fa, foffs, _ = one.flowgen_batch(p)
        fb, _, _ = two.flowgen_batch(p)
        assert np.array_equal(fa, fb)
        _check_sample(oracle, vs, p, f, b, offs, ns, [0, 700, 800, 1535], fb, foffs)
    finally:
        one.close()
        two.close()


def test_random_parameter_fuzz(ctx, vs, oracle):
    """~300 voices with every knob drawn at random inside the reference's accepted ranges (five sampling rates,
    noise on half of them, closure speeds up to 6, DC offsets, random gain/pre-emphasis), both plan kernels,
    chunked and unchunked: flow bit-exact and PCM within 1 LSB of the oracle for EVERY stream"""
    rng = np.random.default_rng(20261018)
    n = 300
    args, seeds = [], []
    for i in range(n):
        fs = int(rng.choice([8000, 11025, 16000, 22051, 44100]))        # -r 22050 is rejected by the reference CLI itself
        f0 = float(rng.uniform(70, 330))
        a = ["-r", str(fs), "-d", f"{rng.uniform(0.5, 1.3):.3f}", "-f", f"{f0:.2f}", "-g", f"{f0 + 20:.2f}"]
        if rng.random() < 0.8:
            a += ["-j", f"{rng.uniform(0, 6):.2f}"]
        if rng.random() < 0.8:
            a += ["-s", f"{rng.uniform(0, 25):.2f}"]
        if rng.random() < 0.5:
            a += ["-n", f"{rng.uniform(3, 45):.1f}"]
        if rng.random() < 0.5:
            a += ["-c", f"{rng.uniform(0.25, 0.95):.3f}"]
        if rng.random() < 0.5:
            a += ["-k", f"{rng.uniform(0.5, 6):.3f}"]
        if rng.random() < 0.4:
            a += ["-z", f"{rng.uniform(0, 1):.3f}"]
        if rng.random() < 0.4:
            a += ["-l", f"{rng.uniform(0, 0.29):.3f}"]
        if rng.random() < 0.5:
            a += ["-a", str(int(rng.integers(500, 18000)))]
        args.append(a)
        seeds.append(int(rng.integers(0, 2**32)))
    p = vs.FlowParams.from_cli(args, seeds)
    # steep closures on very short pitch periods fall outside the defined range (VS_ERANGE): drop those voices
    keep = [i for i in range(n) if vs.flow_validate(p.select(np.array([i])))[0] == vs.VS_OK]
    assert len(keep) > 0.9 * n
    p = p.select(np.array(keep))
    n = p.n
    f = vs.FilterParams(n)
    f.preset[...] = [ord("aiu1234567"[int(k)]) for k in rng.integers(0, 10, n)]
    f.gain[...] = rng.uniform(1.0, 20.0, n).astype(np.float32)
    f.pre[...] = rng.uniform(0.0, 1.0, n).astype(np.float32)
    for warps, chunk in ((0, 0), (1, 0), (-1, 1536)):
        ctx.set_option(vs.OPT_PLAN_WARPS, warps)
        ctx.set_option(vs.OPT_CHUNK_SAMPLES, chunk)
        try:
            flow, foffs, ns = ctx.flowgen_batch(p)
            pcm, offs, _ = ctx.synth_batch(p, f)
        finally:
            ctx.set_option(vs.OPT_PLAN_WARPS, -1)
            ctx.set_option(vs.OPT_CHUNK_SAMPLES, 0)
        _check_sample(oracle, vs, p, f, pcm, offs, ns, range(n), flow, foffs)
    # a batch without any -z takes the pre-multiplied falling-branch tables (one per T2 and K): same streams alone
    idx = np.flatnonzero(p.Kvar == 0)
    assert 0.4 * n < idx.size < n
    p0, f0 = p.select(idx), f.select(idx)
    flow, foffs, ns = ctx.flowgen_batch(p0)
    pcm, offs, _ = ctx.synth_batch(p0, f0)
    _check_sample(oracle, vs, p0, f0, pcm, offs, ns, range(p0.n), flow, foffs)


def test_carry_tolerance_option(ctx, vs, oracle):
    """VS_OPT_CARRY_TOL trades chunk warm-up for carry error: the warm-up of every preset shrinks as the tolerance
    grows, and at 1e-11 (ten times the default) the chunked FP64 waveform is still within 1e-5 of the oracle's and the PCM
    within 1 LSB"""
    n = 20
    args = [f"-d 1.5 -f {90 + 9 * i} -g {130 + 9 * i} -j {0.5 + 0.1 * i:.1f} -s {2 + 0.3 * i:.1f}" for i in range(n)]
    p = vs.FlowParams.from_cli(args, list(range(700, 700 + n)))
    f = vs.FilterParams(n, ("aiu1234567" * 2)[:n])
    want = []
    for i in range(n):
        o_pcm, o_raw = oracle.vowel(oracle.flowgen(_opar(oracle, vs, p, i)), chr(f.preset[i]), gain=float(f.gain[i]), pre=float(f.pre[i]), want_raw=True)
        want.append((o_pcm, o_raw))
    warm, extra = {}, {}
    try:
        for tol in (1e-12, 1e-11, 1e-9):
            ctx.set_option(vs.OPT_CARRY_TOL, tol)
            ctx.set_option(vs.OPT_CHUNK_SAMPLES, 4096)
            warm[tol] = [ctx.filter_warmup(k) for k in "aiu1234567"]
            pcm, offs, ns, raw = ctx.synth_batch(p, f, want_raw=True)
            t = ctx.timing()
            assert t["chunks"] > 3 * n
            extra[tol] = t["warmup_samples"]
            worst = 0.0
            for i in range(n):
                a0, a1 = int(offs[i]), int(offs[i]) + int(ns[i])
                worst = max(worst, float(np.abs(raw[a0:a1] - want[i][1]).max()))
                if tol <= 1e-11:
                    assert int(np.abs(pcm[a0:a1].astype(np.int32) - want[i][0].astype(np.int32)).max()) <= 1, (tol, i)
            print(f"carry tolerance {tol:g}: warm-up {min(warm[tol])}..{max(warm[tol])} samples, {extra[tol]} extra samples, waveform error {worst:.2e}")
            if tol <= 1e-11:
                assert worst <= 1e-5, (tol, worst)
    finally:
        ctx.set_option(vs.OPT_CARRY_TOL, 1e-12)
        ctx.set_option(vs.OPT_CHUNK_SAMPLES, 0)
    assert all(a > b > c for a, b, c in zip(warm[1e-12], warm[1e-11], warm[1e-9]))
    assert extra[1e-12] > extra[1e-11] > extra[1e-9]
