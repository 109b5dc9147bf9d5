"""one flow-only call with glottal noise (cfg3 shard) on device buffers (for ncu): python scripts/prof_noise.py"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np, torch
from voice_synth_b200 import api, workloads

ctx = api.Context()
p, f = workloads.cfg3(n=8192)
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
offs = np.concatenate([[0], np.cumsum(ns)[:-1]]).astype(np.uint64)
for _ in range(2):
    ctx.flowgen_batch(p, out=dev, offsets=offs)
    print(ctx.timing())
