import sys, pathlib, os, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch, numpy as np
from voice_synth_b200 import api, workloads
ctx = api.Context()
p, f = workloads.cfg2()
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
host = torch.empty(int(ns.sum()), dtype=torch.int16).pin_memory().numpy()
for _ in range(3):
    ctx.synth_batch(p, f, out=dev); ctx.sync()
    ctx.synth_batch(p, f, out=host)
os.environ["VS_PROFILE_HOST"] = "1"
print("--- device output", flush=True)
t0 = time.perf_counter(); ctx.synth_batch(p, f, out=dev); t1 = time.perf_counter(); ctx.sync(); t2 = time.perf_counter()
print(f"python call {1e3*(t1-t0):.3f} ms, +sync {1e3*(t2-t1):.3f} ms", ctx.timing(), flush=True)
print("--- pinned host output", flush=True)
t0 = time.perf_counter(); ctx.synth_batch(p, f, out=host); t1 = time.perf_counter()
print(f"python call {1e3*(t1-t0):.3f} ms", ctx.timing(), flush=True)
for slab in (1024, 512, 2048, 4096):
    ctx.set_option(api.OPT_SLAB_STREAMS, slab)
    os.environ.pop("VS_PROFILE_HOST", None)
    ctx.synth_batch(p, f, out=host)
    t0 = time.perf_counter(); ctx.synth_batch(p, f, out=host); t1 = time.perf_counter()
    print(f"slab {slab}: {1e3*(t1-t0):.3f} ms", flush=True)
