"""ad-hoc GPU smoke + timing (not a pytest file): python scripts/gpu_quick.py"""
import sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
from voice_synth_b200 import api, workloads

ctx = api.Context()
p, f = workloads.cfg2()
import torch
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
for name, opts in [("auto", {}), ("nochunk", {1: -1}), ("L2048", {1: 2048}), ("L4096", {1: 4096}), ("L5512", {1: 5512}), ("L8192", {1: 8192}), ("warps1", {5: 1.0}), ("warps3", {5: 3.0})]:
    for k, v in opts.items():
        ctx.set_option(k, v)
    for mode in ("synth", "flow"):
        best = None
        for it in range(4):
            if mode == "synth":
                ctx.synth_batch(p, f, out=dev)
            else:
                ctx.flowgen_batch(p, out=dev)
            t = ctx.timing()
            if best is None or t["render_ms"] < best["render_ms"]:
                best = t
        print(name, mode, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in best.items()},
              "Msamples/s(render)", round(best["samples"] / best["render_ms"] / 1e3, 1), flush=True)
    for k in opts:
        ctx.set_option(k, 0 if k == 1 else 2.0)

# filter-only on device-resident flow: the F-bound rate (producers only copy)
flow = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
ctx.flowgen_batch(p, out=flow)
offs = (torch.arange(p.n, dtype=torch.int64) * int(ns.max())).numpy().astype(np.uint64)
for name, opts in [("auto", {}), ("L5512", {1: 5512}), ("L2756", {1: 2756})]:
    for k, v in opts.items():
        ctx.set_option(k, v)
    best = None
    for it in range(4):
        ctx.vowel_filter_batch(flow, ns, f, out=dev)
        t = ctx.timing()
        if best is None or t["render_ms"] < best["render_ms"]:
            best = t
    print(name, "filter", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in best.items()},
          "Msamples/s(render)", round(best["samples"] / best["render_ms"] / 1e3, 1), flush=True)
    for k in opts:
        ctx.set_option(k, 0)
print("fp64 peak (TFLOP/s, implied MHz):", ctx.fp64_peak())
