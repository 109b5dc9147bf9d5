"""one flow-only call of cfg4 (a single 10-minute stream; 60 s here) for ncu: python scripts/prof_cfg4.py [seconds]"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from voice_synth_b200 import api, workloads

dur = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
ctx = api.Context()
p, f = workloads.cfg4(dur=dur)
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
for _ in range(2):
    ctx.flowgen_batch(p, out=dev)
    print(ctx.timing())
