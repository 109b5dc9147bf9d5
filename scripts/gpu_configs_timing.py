"""timing of BASELINE.json's other configurations on one B200 (device-resident output): python scripts/gpu_configs_timing.py"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np, torch
from voice_synth_b200 import api, workloads

ctx = api.Context()
cases = [("cfg1 single voice 1 s /a/", workloads.cfg1()),
         ("cfg2 4096 x 1 s", workloads.cfg2()),
         ("cfg3 shard 8192 x 2 s, glottal noise 20 dB", workloads.cfg3(n=8192)),
         ("cfg4 one 600 s stream /i/", workloads.cfg4()),
         ("cfg5 slice 131072 x 1 s (1/8 of 1M)", workloads.cfg5(n=131072))]
for name, (p, f) in cases:
    ns = api.flow_nsamples(p)
    dev = torch.empty(int(ns.sum()), dtype=torch.int16, device="cuda")
    offs = np.concatenate([[0], np.cumsum(ns)[:-1]]).astype(np.uint64)
    best = None
    for _ in range(4):
        ctx.synth_batch(p, f, out=dev, offsets=offs)
        t = ctx.timing()
        if best is None or t["total_ms"] < best["total_ms"]:
            best = t
    fbest = None
    for _ in range(3):
        ctx.flowgen_batch(p, out=dev, offsets=offs)
        t = ctx.timing()
        if fbest is None or t["total_ms"] < fbest["total_ms"]:
            fbest = t
    print(f"{name}: synth plan {best['plan_ms']:.3f} + render {best['render_ms']:.3f} ms, {best['chunks']} chunks, "
          f"{best['samples']/best['total_ms']/1e3:.0f} Msamples/s (warm-up {best['warmup_samples']/best['samples']:.2f}x) | "
          f"flow plan {fbest['plan_ms']:.3f} + render {fbest['render_ms']:.3f} ms, {fbest['samples']/fbest['total_ms']/1e3:.0f} Msamples/s", flush=True)
    del dev
