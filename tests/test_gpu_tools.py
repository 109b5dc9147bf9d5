"""The C command-line tools (host/bin) against the fixtures made from the reference tools: same flags,
same PCM payload, same per-period stdout lines.  Needs a GPU (the tools have no CPU path)."""
import hashlib
import os
import pathlib
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parents[1]
BIN = ROOT / "host" / "bin"


def payload(path, hdr=44):
    data = pathlib.Path(path).read_bytes()
    assert data[:4] == b"RIFF" and data[8:16] == b"WAVEfmt " and data[36:40] == b"data"
    return np.frombuffer(data[hdr:], dtype="<i2")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def tools():
    if not (BIN / "flowgen_shimmer").exists():
        subprocess.run(["make", "-s", "-C", str(ROOT), "host"], check=True)
    return BIN


def _vowel_kwargs(extra):
    """-g / -p of a fixture's vowel command line as oracle.vowel() arguments"""
    kw, t = {}, extra.split()
    for flag, key in (("-g", "gain"), ("-p", "pre")):
        if flag in t:
            kw[key] = float(np.float32(float(t[t.index(flag) + 1])))
    return kw


def test_flowgen_and_vowel_tools_match_reference(tools, golden, tmp_path, oracle):
    for c in golden["cases"]:
        if c["name"] not in ("A_cfg1", "B_noise", "F_noise_dc", "O_upper", "J_reject"):
            continue
        env = dict(os.environ, VS_SEED=str(c["seed"]))
        f = tmp_path / "f.wav"
        r = subprocess.run([str(tools / "flowgen_shimmer"), "-o", str(f)] + c["args"].split(), env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        pcm = payload(f)
        assert pcm.size == c["n"] and sha(pcm) == c["sha256"], c["name"]
        # per-period lines: shimmer draws "%5.2f \n" and "SNRdb = %5.2f" exactly as the reference prints them
        S = re.findall(r"^\s*(-?\d+\.\d\d) $", r.stdout, flags=re.M)
        snr = re.findall(r"SNRdb = \s*(-?[\d.]+|-?nan|-?inf)", r.stdout)
        assert len(snr) == c["snr_count"] and snr[:8] == c["snr_head"], c["name"]
        assert len(S) in (c["S_count"], c["S_count"] + 1), c["name"]           # the reference glues its first line to "Wait..."
        assert r.stdout.rstrip().endswith("done")
        for v in c["vowels"]:
            o = tmp_path / "o.wav"
            r2 = subprocess.run([str(tools / "vowel"), "-i", str(f), "-o", str(o), "-v", v["preset"]] + v["extra"].split(),
                                env=env, capture_output=True, text=True)
            assert r2.returncode == 0, r2.stderr
            out = payload(o)
            assert out.size == c["n"]
            assert out[:16].tolist() == v["head"]
            if "-n" in v["extra"].split():
                # `vowel -n` filters in exact mode and runs vs_vowel_noise_batch with the reference's seed: bit-exact
                assert sha(out) == v["sha256"], (c["name"], v)
            else:
                # the fast filter's contract is +-1 LSB (it usually lands on the very same int16)
                want = oracle.vowel(pcm, v["preset"], **_vowel_kwargs(v["extra"]))
                assert int(np.abs(out.astype(np.int32) - want.astype(np.int32)).max()) <= 1, (c["name"], v)
                assert sha(want) == v["sha256"], (c["name"], v)          # ... and the oracle is the reference here


def test_tools_reject_what_the_reference_rejects(tools, tmp_path):
    for bad in (["-d", "0.4"], ["-f", "130"], ["-a", "32767"], ["-q", "1"]):
        r = subprocess.run([str(tools / "flowgen_shimmer"), "-o", str(tmp_path / "x.wav")] + bad, capture_output=True, text=True)
        assert r.returncode == 0 and "usage" in r.stdout
    r = subprocess.run([str(tools / "vowel"), "-i", "nope.wav", "-o", str(tmp_path / "y.wav"), "-v", "A"], capture_output=True, text=True)
    assert "usage" in r.stdout


def test_vowel_tool_reads_the_reference_lp64_header(tools, tmp_path):
    """a WAV with the 72-byte header of the reference's 64-bit build is accepted"""
    z = np.load(ROOT / "tests" / "golden" / "cfg1_seed42.npz")
    hdr = bytearray(72)
    hdr[0:4] = b"RIFF"; hdr[16:20] = b"WAVE"; hdr[20:24] = b"fmt "; hdr[60:64] = b"data"
    hdr[24] = 16; hdr[32] = 1; hdr[34] = 1
    hdr[40:44] = (22050).to_bytes(4, "little"); hdr[48:52] = (44100).to_bytes(4, "little"); hdr[56] = 2; hdr[58] = 16
    hdr[64:68] = (44100).to_bytes(4, "little")
    f = tmp_path / "ref.wav"
    f.write_bytes(bytes(hdr) + z["flow"].astype("<i2").tobytes())
    o = tmp_path / "o.wav"
    r = subprocess.run([str(tools / "vowel"), "-i", str(f), "-o", str(o), "-v", "a"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(payload(o), z["vowel_a"])


def test_batch_driver(tools, golden, tmp_path, oracle):
    cases = [c for c in golden["cases"] if c["vowels"]][:8]
    lines = []
    for i, c in enumerate(cases):
        lines.append(f"{tmp_path}/v{i}.wav {c['vowels'][0]['preset']} {c['seed']} {c['args']}")
    m = tmp_path / "manifest.txt"
    m.write_text("# out vowel seed flags\n" + "\n".join(lines) + "\n")
    r = subprocess.run([str(tools / "vs_batch"), "-m", str(m)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    for i, c in enumerate(cases):
        v = c["vowels"][0]
        if v["extra"]:
            continue
        out = payload(tmp_path / f"v{i}.wav")
        par = oracle.flow_par_from_cli(["-o", "x"] + c["args"].split(), c["seed"])
        want = oracle.vowel(oracle.flowgen(par), v["preset"])
        assert out.size == want.size
        assert np.abs(out.astype(np.int32) - want.astype(np.int32)).max() <= 1, c["name"]


def test_batch_driver_slabs_and_writer_pool(tools, golden, tmp_path):
    """N2: the manifest cut into slabs that alternate between two pinned buffers while writer threads put the
    previous slab on disk -- same files as one slab, one writer"""
    cases = [c for c in golden["cases"] if c["vowels"]]
    lines = []
    for rep in range(3):
        for i, c in enumerate(cases):
            lines.append(f"{{d}}/r{rep}_v{i}.wav {c['vowels'][0]['preset']} {c['seed'] + rep} {c['args']}")
    outs = []
    for name, extra in (("one", ["-S", "100000", "-W", "1"]), ("slabs", ["-S", "5", "-W", "4"])):
        d = tmp_path / name
        d.mkdir()
        m = tmp_path / f"{name}.txt"
        m.write_text("\n".join(lines).format(d=d) + "\n")
        r = subprocess.run([str(tools / "vs_batch"), "-m", str(m)] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        outs.append({f.name: f.read_bytes() for f in sorted(d.iterdir())})
    assert len(outs[0]) == len(lines)
    assert outs[0] == outs[1]
    # unwritable output path: reported, non-zero exit
    m = tmp_path / "bad.txt"
    m.write_text(f"{tmp_path}/no_such_dir/x.wav a 1 -d 0.5\n")
    r = subprocess.run([str(tools / "vs_batch"), "-m", str(m)], capture_output=True, text=True)
    assert r.returncode != 0


def test_acoustic_tool_measures_what_flowgen_made(tools, tmp_path):
    """host/bin/acoustic on WAV files written by host/bin/flowgen_shimmer: its line per file carries the
    statistics tests/analysis_ref.py computes from the same samples"""
    import analysis_ref as ar
    files = []
    for k, args in enumerate(["-d 1 -f 110 -g 140 -j 2 -s 6", "-d 0.8 -f 201 -g 260", "-d 1.2 -f 87 -g 120 -j 0.5 -a 9000"]):
        f = tmp_path / f"v{k}.wav"
        r = subprocess.run([str(tools / "flowgen_shimmer"), "-o", str(f)] + args.split(), env=dict(os.environ, VS_SEED=str(100 + k)),
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        files.append(f)
    r = subprocess.run([str(tools / "acoustic")] + [str(f) for f in files], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == len(files)
    for f, line in zip(files, lines):
        t = line.split()
        want = ar.stats(payload(f), fs=22050)
        assert t[0] == str(f) and int(t[1]) == 22050 and int(t[2]) == payload(f).size and int(t[3]) == want["cycles"]
        for got, key, nd in zip(t[4:9], ("f0_hz", "jitter_pct", "shimmer_pct", "mean_period", "mean_peak"), (3, 4, 4, 3, 2)):
            assert abs(float(got) - want[key]) <= 0.6 * 10 ** -nd + 1e-6 * abs(want[key]), (f.name, key, got, want[key])
    # the second voice has neither jitter nor shimmer: its cycles are all alike
    assert float(lines[1].split()[5]) == 0.0 and float(lines[1].split()[6]) == 0.0
    assert "usage" in subprocess.run([str(tools / "acoustic"), "-l", "5", "-h", "2", str(files[0])], capture_output=True, text=True).stdout
