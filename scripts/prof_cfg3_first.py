"""host time of the first calls of a cfg3 shard (allocation / first-touch effects): python scripts/prof_cfg3_first.py"""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from voice_synth_b200 import api, workloads

stream = torch.cuda.current_stream()
ctx = api.Context(devices=[0], stream=stream.cuda_stream)
p3, f3 = workloads.cfg3(n=8192, first=0)
n3 = int(api.flow_nsamples(p3).sum())
out3 = torch.empty(n3, dtype=torch.int16, device="cuda")
torch.cuda.synchronize()
for k in range(16):
    t0 = time.perf_counter()
    ctx.synth_batch(p3, f3, out=out3)
    t1 = time.perf_counter()
    if k in (2, 12):
        ctx.sync()
    t2 = time.perf_counter()
    print(f"call {k}: enqueue {1e3*(t1-t0):.3f} ms, sync {1e3*(t2-t1):.3f} ms")
