"""pure F-phase rate: one wave of unchunked rows through the filter (device-resident random flow)"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np, torch
from voice_synth_b200 import api
ctx = api.Context()
ctx.set_option(api.OPT_CHUNK_SAMPLES, -1)
for nrows, nsamp in ((18944, 4032), (18944, 8064), (9472, 8064), (37888, 4032)):
    f = api.FilterParams(nrows, "a")
    flow = torch.randint(-2000, 12000, (nrows * nsamp,), dtype=torch.int16, device="cuda")
    out = torch.zeros_like(flow)
    ns = np.full(nrows, nsamp, dtype=np.uint64)
    best = 1e9
    for _ in range(5):
        ctx.vowel_filter_batch(flow, ns, f, out=out)
        t = ctx.timing()
        best = min(best, t["render_ms"])
    cyc_per_sample = best * 1e-3 * 1.9e9 / (nsamp * max(1, nrows / 18944))
    print(f"rows {nrows} x {nsamp}: {best:.4f} ms -> {nrows*nsamp/best/1e3:.0f} Msamples/s, ~{cyc_per_sample:.1f} cycles per sample-step per consumer (at 1.9 GHz), DFMA/s {25*nrows*nsamp/best/1e9:.2f} T")
