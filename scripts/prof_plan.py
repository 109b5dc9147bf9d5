"""plan-kernel timing, one thread vs one warp per stream, over batch sizes: python scripts/prof_plan.py"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np, torch
from voice_synth_b200 import api, workloads

ctx = api.Context()
for name, make in (("cfg2", workloads.cfg2), ("cfg5", workloads.cfg5), ("cfg3", workloads.cfg3)):
    for n in (256, 1024, 4096, 16384, 65536):
        p, f = make(n=n)
        ns = api.flow_nsamples(p)
        dev = torch.empty(int(ns.sum()), dtype=torch.int16, device="cuda")
        offs = np.concatenate([[0], np.cumsum(ns)[:-1]]).astype(np.uint64)
        res = []
        for w in (0, 1):
            ctx.set_option(api.OPT_PLAN_WARPS, w)
            best = 1e9
            for _ in range(3):
                ctx.synth_batch(p, f, out=dev, offsets=offs)
                best = min(best, ctx.timing()["plan_ms"])
            res.append(best)
        ctx.set_option(api.OPT_PLAN_WARPS, -1)
        print(f"{name} n={n}: plan thread/stream {res[0]:.3f} ms, warp/stream {res[1]:.3f} ms", flush=True)
        del dev
