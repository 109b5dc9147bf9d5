"""how long do K pipelined calls of the bench batch take from an idle device: total = ramp + K x step
python scripts/prof_ramp.py"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from voice_synth_b200 import api, workloads

stream = torch.cuda.current_stream()
ctx = api.Context(devices=[0], stream=stream.cuda_stream)
p, f = workloads.cfg2()
n = int(api.flow_nsamples(p).sum())
dev = torch.empty(n, dtype=torch.int16, device="cuda")
for _ in range(8):
    ctx.synth_batch(p, f, out=dev)
ctx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
prev = None
for K in (1, 2, 3, 4, 5, 6, 8, 10, 20, 40, 80):
    best = 1e9
    for rep in range(5):
        ctx.sync()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(K):
            ctx.synth_batch(p, f, out=dev)
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"K={K:3d}: {best:.3f} ms total, {best / K:.4f} per step" + (f", +{(best - prev[1]) / (K - prev[0]):.4f} per extra step" if prev else ""))
    prev = (K, best)
