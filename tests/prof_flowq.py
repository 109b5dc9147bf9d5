import sys, pathlib
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from voice_synth_b200 import api, workloads
ctx = api.Context()
p, f = workloads.cfg2()
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
for name, opts in [("auto", {}), ("L384", {1: 384}), ("L576", {1: 576}), ("L768", {1: 768}), ("L1536", {1: 1536})]:
    for k, v in opts.items(): ctx.set_option(k, v)
    best = None
    for it in range(5):
        ctx.flowgen_batch(p, out=dev)
        t = ctx.timing()
        if best is None or t["render_ms"] < best["render_ms"]: best = t
    print(name, "flow render_ms", round(best["render_ms"], 4), "chunks", best["chunks"], "GB/s", round(best["samples"] * 2 / best["render_ms"] / 1e6, 1), flush=True)
    for k in opts: ctx.set_option(k, 0)
