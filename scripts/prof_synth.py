"""one fused-synthesis call of the bench workload on device buffers (for ncu): python scripts/prof_synth.py [n_streams] [chunk]"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from voice_synth_b200 import api, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = api.Context()
if len(sys.argv) > 2:
    ctx.set_option(api.OPT_CHUNK_SAMPLES, float(sys.argv[2]))
p, f = workloads.cfg2(n=n)
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
for _ in range(3):
    ctx.synth_batch(p, f, out=dev)
    print(ctx.timing())
