"""five pipelined steps of the bench workload between cudaProfilerStart/Stop, for
   ncu --replay-mode application-range (DRAM bytes of whole steps, kernels overlapping as they do in the bench):
   python scripts/prof_range.py [steps]"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from voice_synth_b200 import api, workloads

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ctx = api.Context()
p, f = workloads.cfg2()
ns = api.flow_nsamples(p)
dev = torch.zeros(int(ns.sum()), dtype=torch.int16, device="cuda")
for _ in range(6):
    ctx.synth_batch(p, f, out=dev)
ctx.sync()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(steps):
    ctx.synth_batch(p, f, out=dev)
ctx.sync()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("steps", steps, "samples per step", int(ns.sum()), ctx.timing())
