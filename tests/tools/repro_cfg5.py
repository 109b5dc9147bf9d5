import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[2]))
import os
import numpy as np, torch
from voice_synth_b200 import api, workloads
n = int(sys.argv[1]); dur = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0; first = int(os.environ.get('FIRST','0'))
ctx = api.Context()
if os.environ.get('CHUNK'): ctx.set_option(api.OPT_CHUNK_SAMPLES, float(os.environ['CHUNK']))
mode = os.environ.get('MODE', 'both')
p, f = workloads.cfg5(n=n, first=first, dur=dur)
if os.environ.get('PRESET'): f.preset[...] = ord(os.environ['PRESET'])
if os.environ.get('NONOISE'):
    p.flags[...] &= 3; p.DC[...] = 0
ns = api.flow_nsamples(p)
dev = torch.empty(int(ns.sum()), dtype=torch.int16, device="cuda")
offs = np.concatenate([[0], np.cumsum(ns)[:-1]]).astype(np.uint64)
if mode in ('both','synth'):
    ctx.synth_batch(p, f, out=dev, offsets=offs); print("synth", ctx.timing())
if mode in ('both','flow'):
    ctx.flowgen_batch(p, out=dev, offsets=offs); print("flow", ctx.timing())
