#!/usr/bin/env python
"""Summarise a k_forward trace (-DPERSIST_TRACE=1 build, GF2B200_TRACE_FILE=...):
per-panel phase times (us after the panel's first CTA started), averaged over panel ranges.
    python scripts/trace_forward.py file"""
import sys
import warnings
import numpy as np
warnings.simplefilter("ignore")
raw = np.fromfile(sys.argv[1], dtype=np.uint64)
nw, G = int(raw[0]), int(raw[1])
tp = raw[2:2 + nw + 2].astype(np.int64)
tr = raw[2 + nw + 2:].astype(np.int64).reshape(nw, G, 8)
def rng(a, b):
    sl = tr[a:b]
    t0 = sl[:, :, 0].min(axis=1)
    def rel(slot, fn):
        v = sl[:, :, slot].astype(float); v[v == 0] = np.nan
        return np.nanmean(fn(v - t0[:, None], axis=1)) / 1e3
    panel = np.diff(tp[a:b + 1]).astype(float)
    return {"panel": np.nanmean(panel) / 1e3, "start_skew": rel(0, np.nanmax), "tables_mean": rel(1, np.nanmean),
            "units_mean": rel(2, np.nanmean), "units_max": rel(2, np.nanmax), "flag_max": rel(3, np.nanmax),
            "apply_mean": rel(4, np.nanmean), "apply_max": rel(4, np.nanmax), "barrier_max": rel(5, np.nanmax),
            "srch_start": rel(6, np.nanmax), "srch_end": rel(7, np.nanmax)}
# the search CTA (the one with a non-zero slot 6) re-uses its slots 1 / 0: window search done / finalize done
srch = tr[:, :, 6] > 0
def srch_phase(a, b):
    sl = tr[a:b]; m_ = srch[a:b]
    if not m_.any(): return None
    t6 = sl[:, :, 6][m_].astype(float); t1 = sl[:, :, 1][m_].astype(float); t0 = sl[:, :, 0][m_].astype(float); t7 = sl[:, :, 7][m_].astype(float)
    return {"window_search": np.mean(t1 - t6) / 1e3, "finalize": np.mean(t0 - t1) / 1e3, "fence+release+sync": np.mean(t7 - t0) / 1e3}
step = max(1, nw // 8)
for a in range(0, nw - 1, step):
    b = min(nw - 1, a + step)
    o = rng(a, b)
    print(f"panels {a:5d}-{b:5d}: " + " ".join(f"{k}={v:7.2f}" for k, v in o.items()))
    sp = srch_phase(a, b)
    if sp: print("      search CTA: " + " ".join(f"{k}={v:6.2f}" for k, v in sp.items()))
