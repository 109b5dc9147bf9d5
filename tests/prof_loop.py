import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch, numpy as np
from voice_synth_b200 import api, workloads
ctx = api.Context()
p, f = workloads.cfg2()
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
for _ in range(5): ctx.synth_batch(p, f, out=dev)
ctx.sync()
N = 200
ts = []
t0 = time.perf_counter()
for _ in range(N):
    a = time.perf_counter(); ctx.synth_batch(p, f, out=dev); ts.append(time.perf_counter() - a)
t1 = time.perf_counter(); ctx.sync(); t2 = time.perf_counter()
ts = np.array(ts) * 1e3
print(f"host per call: median {np.median(ts):.3f} ms, mean {ts.mean():.3f}, max {ts.max():.3f}; loop {1e3*(t1-t0)/N:.3f} ms/step, +final sync {1e3*(t2-t1):.3f} ms")
# same loop with a sync every step (no overlap)
t0 = time.perf_counter()
for _ in range(50):
    ctx.synth_batch(p, f, out=dev); ctx.sync()
print(f"synced loop: {1e3*(time.perf_counter()-t0)/50:.3f} ms/step")
