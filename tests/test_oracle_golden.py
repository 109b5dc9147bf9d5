"""The oracle (oracle/vs_oracle.c) against the committed fixtures made from the unmodified reference
binaries (tests/golden/make_golden.py).  CPU only; this is what pins the oracle."""
import hashlib

import numpy as np
import pytest


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_rng_known_answers(golden, oracle):
    # glibc random()/srandom() TYPE_3; vectors were cross-checked against libc itself when generated
    for seed, want in golden["rng"].items():
        assert oracle.random_sequence(int(seed), len(want)) == want
    # SURVEY.md 8c KATs, literal
    assert oracle.random_sequence(1, 3) == [1804289383, 846930886, 1681692777]
    assert oracle.random_sequence(42, 3) == [71876166, 708592740, 1483128881]


def _cases(golden):
    return golden["cases"]


def test_flow_pcm_matches_reference(golden, oracle):
    for c in _cases(golden):
        par = oracle.flow_par_from_cli(["-o", "x.wav"] + c["args"].split(), c["seed"])
        assert par is not None, c["name"]
        pcm = oracle.flowgen(par)
        assert pcm.size == c["n"], c["name"]
        assert pcm[:16].tolist() == c["head"], c["name"]
        assert int(pcm.astype(np.int64).sum()) == c["sum"], c["name"]
        assert sha(pcm) == c["sha256"], (c["name"], c["seed"])


def test_vowel_pcm_matches_reference(golden, oracle):
    for c in _cases(golden):
        if not c["vowels"]:
            continue
        par = oracle.flow_par_from_cli(["-o", "x.wav"] + c["args"].split(), c["seed"])
        flow = oracle.flowgen(par)
        fs = int(par.fs)
        for v in c["vowels"]:
            ex = v["extra"].split()
            opt = {ex[i]: float(ex[i + 1]) for i in range(0, len(ex), 2)}
            gain, pre = opt.get("-g", 10.0), opt.get("-p", 1.0)
            if "-n" in opt:
                out = oracle.vowel_noise(flow, v["preset"], opt["-n"], c["seed"], gain=gain, pre=pre, fs=fs)
            else:
                out = oracle.vowel(flow, v["preset"], gain=gain, pre=pre)
            assert out[:16].tolist() == v["head"], (c["name"], v)
            assert int((np.abs(out) == 32767).sum()) == v["clipped"], (c["name"], v)
            assert sha(out) == v["sha256"], (c["name"], c["seed"], v["preset"], v["extra"])


def test_cfg1_full_vectors(oracle):
    import pathlib
    z = np.load(pathlib.Path(__file__).parent / "golden" / "cfg1_seed42.npz")
    par = oracle.flow_par_from_cli("-o x.wav -d 1 -f 120 -j 1 -s 3".split(), 42)
    flow, log = oracle.flowgen(par, want_log=True)
    assert np.array_equal(flow, z["flow"])
    assert np.array_equal(oracle.vowel(flow, "a"), z["vowel_a"])
    # SURVEY.md 8c period KATs, seed 42
    assert len(log) == 119 and int(log["ndraws"].sum()) == 357
    assert [(int(r["T"]), int(r["T2"]), int(r["T3"])) for r in log[:5]] == [(180, 51, 86), (180, 51, 86), (181, 51, 86), (177, 51, 86), (176, 51, 86)]
    assert [float(r["A"]) for r in log[:5]] == [11757.62109375, 11350.3642578125, 11856.6796875, 11666.134765625, 11167.0849609375]
    assert float(log["Knew"][0]) == 0.6499999761581421


def test_period_log_consistency(oracle):
    par = oracle.flow_par_from_cli("-o x.wav -d 2 -f 120 -j 1 -s 3 -n 20".split(), 42)
    flow, log = oracle.flowgen(par, want_log=True)
    assert len(log) == 257 and int(log["ndraws"].sum()) == 22809
    assert [(int(r["T"]), int(r["T4"]), int(r["ndraws"])) for r in log[:3]] == [(180, 0, 97), (179, 0, 96), (176, 0, 93)]
    starts = np.concatenate([[0], np.cumsum(log["T"])[:-1]])
    assert np.array_equal(starts, log["start"])
    assert starts[-1] < flow.size <= starts[-1] + log["T"][-1]


@pytest.mark.parametrize("args", ["-o x -d 0.4", "-o x -j 1001", "-o x -f 130", "-o x -r 22050", "-d 1",
                                  "-o x -a 32767", "-o x -l 0.31", "-o x -q 1", "-o", "-o x -n 51", "-o x -s 101"])
def test_cli_rejections(oracle, args):
    # every one of these makes the reference print usage() and exit(0) (flowgen_shimmer.c:128-219,463-547)
    assert oracle.flow_par_from_cli(args.split(), 1) is None


def test_cli_conversions(oracle):
    p = oracle.flow_par_from_cli("-O x -J 1 -S 3 -N 20 -R 16000 -L 0.1 -A 10000".split(), 1)
    assert p.jitter == np.float32(0.01) and p.shimmer == np.float32(np.float32(3) / np.float32(100))
    assert p.noise == np.float32(100.0) and p.fs == 16000 and p.DC == np.float32(np.float32(0.1) * np.float32(10000))
    p = oracle.flow_par_from_cli("-o x -n 20".split(), 1)
    assert p.DC == np.float32(0.25) and p.has_noise == 1 and p.has_jitter == 0


def test_round2int_semantics(oracle):
    r = oracle.lib().vso_round2int
    # round-half-DOWN, symmetric clip, -32768 never produced (vowel_new.c:413-427)
    assert [r(x) for x in (0.5, 1.5, 2.5, -0.5, -1.5, 0.5000001, -0.4999999)] == [0, 1, 2, -1, -2, 1, 0]
    assert [r(x) for x in (40000.0, -40000.0, 32767.4, -32767.6, 1e300, -1e300)] == [32767, -32767, 32767, -32767, 32767, -32767]
