"""filter-only call on device-resident flow (for ncu): python scripts/prof_filter.py"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np, torch
from voice_synth_b200 import api, workloads
ctx = api.Context()
p, f = workloads.cfg2()
f.preset[...] = ord("a")          # one preset -> one launch
ns = api.flow_nsamples(p)
flow = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
ctx.flowgen_batch(p, out=flow)
for _ in range(3):
    ctx.vowel_filter_batch(flow, ns, f, out=dev)
    print(ctx.timing())
