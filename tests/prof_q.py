import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from voice_synth_b200 import api, workloads
ctx = api.Context()
p, f = workloads.cfg2()
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
best = None
for it in range(6):
    ctx.synth_batch(p, f, out=dev)
    t = ctx.timing()
    if best is None or t["render_ms"] < best["render_ms"]: best = t
print("synth render_ms", round(best["render_ms"], 4), "chunks", best["chunks"])
best = None
for it in range(4):
    ctx.flowgen_batch(p, out=dev)
    t = ctx.timing()
    if best is None or t["render_ms"] < best["render_ms"]: best = t
print("flow render_ms", round(best["render_ms"], 4))
