"""cfg5 sweep anatomy (16 384 new utterances per call, two pinned buffers): python scripts/prof_cfg5_host.py [same]
`same`: every call gets the first part's parameters again (the host then skips descriptor building)"""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from voice_synth_b200 import api, workloads

same = "same" in sys.argv[1:]
ctx = api.Context()
sub, nsub = 16384, 8
parts = [workloads.cfg5(n=sub, first=k * sub) for k in range(nsub)]
if same:
    parts = [parts[0]] * nsub
n5 = [int(api.flow_nsamples(pp).sum()) for pp, _ in parts]
ring = [torch.empty(max(n5), dtype=torch.int16).pin_memory().numpy() for _ in range(2)]
ctx.set_option(api.OPT_ASYNC_HOST, 1)
for rep in range(3):
    t0 = time.perf_counter()
    marks = []
    for k, (pp, ff) in enumerate(parts):
        ctx.synth_batch(pp, ff, out=ring[k & 1])
        marks.append(time.perf_counter() - t0)
    ctx.sync()
    tot = time.perf_counter() - t0
    print(f"rep {rep} same={same}: calls returned at " + " ".join(f"{1e3 * m:.1f}" for m in marks) + f" ms; sweep {1e3 * tot:.1f} ms, {sum(n5) * 2 / tot / 1e9:.1f} GB/s", flush=True)
# the same bytes as bare copies
dev = torch.empty(max(n5), dtype=torch.int16, device="cuda")
host = torch.from_numpy(ring[0])
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(nsub):
    host.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(f"bare copies: {1e3 * tot:.1f} ms, {max(n5) * nsub * 2 / tot / 1e9:.1f} GB/s")
