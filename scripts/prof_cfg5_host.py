"""host-side cost of the cfg5 calls (16 384 new utterances per call): VS_PROFILE_HOST=1 python scripts/prof_cfg5_host.py"""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from voice_synth_b200 import api, workloads

ctx = api.Context()
sub, nsub = 16384, 4
parts = [workloads.cfg5(n=sub, first=k * sub) for k in range(nsub)]
n5 = [int(api.flow_nsamples(pp).sum()) for pp, _ in parts]
ring = [torch.empty(max(n5), dtype=torch.int16).pin_memory().numpy() for _ in range(2)]
ctx.set_option(api.OPT_ASYNC_HOST, 1)
for rep in range(2):
    t0 = time.perf_counter()
    for k, (pp, ff) in enumerate(parts):
        t1 = time.perf_counter()
        ctx.synth_batch(pp, ff, out=ring[k & 1])
        print(f"rep {rep} call {k}: host {1e3 * (time.perf_counter() - t1):.2f} ms", flush=True)
    ctx.sync()
    print(f"rep {rep}: sweep {1e3 * (time.perf_counter() - t0):.1f} ms for {sum(n5) * 2 / 1e9:.2f} GB", flush=True)
