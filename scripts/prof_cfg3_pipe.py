"""cfg3 shard, pipelined calls on a device buffer (what bench.py's other_workloads.cfg3 times), optionally after the
things bench.py does first: python scripts/prof_cfg3_pipe.py [stream] [dev] [host] [flow] [filt]"""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from voice_synth_b200 import api, workloads

what = set(sys.argv[1:])
stream = torch.cuda.current_stream()
ctx = api.Context(devices=[0], stream=stream.cuda_stream) if "stream" in what else api.Context()
p, f = workloads.cfg2()
ns = api.flow_nsamples(p)
n2 = int(ns.sum())
dev = torch.zeros(n2, dtype=torch.int16, device="cuda")
if "dev" in what:
    for _ in range(30):
        ctx.synth_batch(p, f, out=dev)
    ctx.sync()
if "host" in what:
    host = torch.empty(n2, dtype=torch.int16).pin_memory().numpy()
    ctx.set_option(api.OPT_ASYNC_HOST, 1)
    for _ in range(10):
        ctx.synth_batch(p, f, out=host)
    ctx.sync()
    ctx.set_option(api.OPT_ASYNC_HOST, 0)
flow_dev = torch.zeros(n2, dtype=torch.int16, device="cuda")
if "flow" in what:
    for _ in range(6):
        ctx.flowgen_batch(p, out=flow_dev)
        ctx.timing()
if "filt" in what:
    for _ in range(6):
        ctx.vowel_filter_batch(flow_dev, ns, f, out=dev)
        ctx.timing()
del flow_dev
p3, f3 = workloads.cfg3(n=8192, first=0)
n3 = int(api.flow_nsamples(p3).sum())
out3 = torch.empty(n3, dtype=torch.int16, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    for _ in range(3):
        ctx.synth_batch(p3, f3, out=out3)
    ctx.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(10):
        ctx.synth_batch(p3, f3, out=out3)
    e1.record(stream)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    ctx.sync()
    t2 = time.perf_counter()
    t = ctx.timing()
    print(f"{sorted(what)} rep {rep}: enqueue {1e3*(t1-t0)/10:.3f} ms/call, total {1e3*(t2-t0)/10:.3f} ms/call, events {e0.elapsed_time(e1)/10:.3f}, plan {t['plan_ms']:.3f} render {t['render_ms']:.3f} chunks {t['chunks']}")
