"""the bench loop (same call, device output, no sync in between) with host-side section times: python scripts/prof_loop.py"""
import os, sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from voice_synth_b200 import api, workloads

prof = len(sys.argv) > 1
if prof:
    os.environ["VS_PROFILE_HOST"] = "1"
ctx = api.Context()
p, f = workloads.cfg2()
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
for _ in range(5):
    ctx.synth_batch(p, f, out=dev)
ctx.sync()
N = 6 if prof else 200
t0 = time.perf_counter()
for _ in range(N):
    ctx.synth_batch(p, f, out=dev)
t1 = time.perf_counter()
ctx.sync()
t2 = time.perf_counter()
print(f"enqueue {1e3*(t1-t0)/N:.4f} ms/call, with final sync {1e3*(t2-t0)/N:.4f} ms/call", ctx.timing())
# python-side share: the ctypes argument marshalling alone
t0 = time.perf_counter()
for _ in range(200):
    p._c(); f._c()
t1 = time.perf_counter()
print(f"python marshalling {1e3*(t1-t0)/200:.4f} ms/call")
