"""plan kernel time (thread form) for the bench batch on an otherwise idle GPU: python scripts/prof_plan_nt.py"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from voice_synth_b200 import api, workloads
ctx = api.Context()
ctx.set_option(api.OPT_PLAN_WARPS, 0)
p, f = workloads.cfg2()
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
best = 1e9
for _ in range(5):
    ctx.synth_batch(p, f, out=dev); ctx.sync()
    best = min(best, ctx.timing()["plan_ms"])
print("thread-form plan_ms", round(best, 4))
