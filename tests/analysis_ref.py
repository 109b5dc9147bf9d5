"""numpy restatement of vs_flow_analyze_batch (include/voicesynth.h): the definitions are the library's own
(the reference's `acoustic` tools are not in its tree), this file states them a second time, independently of the CUDA
kernels, for the tests."""
import numpy as np


def onsets(x, lo=0, hi=0):
    """two-threshold trigger: armed at the start and by x <= lo, fires at the first x > hi while armed"""
    x = np.asarray(x, dtype=np.int32)
    arm, dis = x <= lo, x > hi
    # state in front of sample m = kind of the last arm/disarm before m (armed if none)
    kind = np.where(arm, 1, np.where(dis, -1, 0))
    idx = np.where(kind != 0, np.arange(len(x)), -1)
    last = np.maximum.accumulate(idx)
    prev = np.concatenate([[-1], last[:-1]])
    armed_before = np.where(prev < 0, True, kind[np.maximum(prev, 0)] == 1)
    return np.nonzero(dis & armed_before)[0]


def stats(x, fs=22050, lo=0, hi=0):
    o = onsets(x, lo, hi)
    out = {"onsets": len(o), "cycles": 0, "flags": 0, "f0_hz": 0.0, "jitter_pct": 0.0, "shimmer_pct": 0.0, "mean_period": 0.0, "mean_peak": 0.0}
    if len(o) > len(x) // 16 + 4:
        out["flags"] = 1
        return out
    nc = max(len(o) - 1, 0)
    out["cycles"] = nc
    if nc >= 1:
        T = np.diff(o).astype(np.int64)
        xi = np.asarray(x, dtype=np.int64)
        P = np.array([xi[o[k]: o[k + 1]].max() for k in range(nc)], dtype=np.int64)
        mT, mP = T.sum() / nc, P.sum() / nc
        out["mean_period"], out["mean_peak"], out["f0_hz"] = mT, mP, fs / mT
        if nc >= 2:
            out["jitter_pct"] = 100.0 * (np.abs(np.diff(T)).sum() / (nc - 1)) / mT
            out["shimmer_pct"] = 100.0 * (np.abs(np.diff(P)).sum() / (nc - 1)) / mP if mP != 0 else 0.0
    return out
