"""A/B timing of debug switches (option 100) on the bench workload: python scripts/prof_ab.py"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from voice_synth_b200 import api, workloads

ctx = api.Context()
p, f = workloads.cfg2()
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
for dbg in (0, 1):
    ctx.set_option(100, dbg)
    for mode in ("synth", "flow"):
        best = 1e9
        for _ in range(5):
            (ctx.synth_batch(p, f, out=dev) if mode == "synth" else ctx.flowgen_batch(p, out=dev))
            best = min(best, ctx.timing()["render_ms"])
        print("debug", dbg, mode, round(best, 4), flush=True)
