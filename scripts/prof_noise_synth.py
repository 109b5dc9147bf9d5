"""one fused-synthesis call of a cfg3 shard (8192 x 2 s, glottal noise) on device buffers (for ncu): python scripts/prof_noise_synth.py"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np, torch
from voice_synth_b200 import api, workloads

ctx = api.Context()
p, f = workloads.cfg3(n=8192)
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
for _ in range(3):
    ctx.synth_batch(p, f, out=dev)
    print(ctx.timing())
