"""one flow-only call of the bench workload on device buffers (for ncu): python scripts/prof_flow.py [chunk]"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from voice_synth_b200 import api, workloads

ctx = api.Context()
if len(sys.argv) > 1:
    ctx.set_option(api.OPT_CHUNK_SAMPLES, float(sys.argv[1]))
p, f = workloads.cfg2()
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
for _ in range(3):
    ctx.flowgen_batch(p, out=dev)
    print(ctx.timing())
