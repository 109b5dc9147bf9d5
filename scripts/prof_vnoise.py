"""vowel -n on the bench batch (4096 x 22050 samples, device-resident): python scripts/prof_vnoise.py"""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np, torch
from voice_synth_b200 import api, workloads

ctx = api.Context()
p, f = workloads.cfg2()
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
ctx.synth_batch(p, f, out=dev)
ctx.sync()
offs = np.concatenate([[0], np.cumsum(ns)[:-1]]).astype(np.uint64)
snr = np.full(p.n, 100.0, dtype=np.float32)
seeds = (1000 + np.arange(p.n)).astype(np.uint32)
nsu = np.ascontiguousarray(ns, dtype=np.uint64)
for it in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = ctx.L.vs_vowel_noise_batch(ctx.h, api._ptr(dev), api._ptr(offs), nsu.ctypes.data, snr.ctypes.data, None, seeds.ctypes.data, p.n)
    assert rc == 0
    print(f"vs_vowel_noise_batch (C call, synchronous), {p.n} streams x {int(ns[0])} samples: {(time.perf_counter() - t0) * 1e3:.3f} ms", flush=True)
