import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from libpointmatcher_b200 import capi, synth
rd, rf, T = synth.scan_pair(1_000_000)
with capi.Context(0) as ctx:
    ctx.set_reference(rf); ctx.set_reading(rd)
    ctx.knn(None, 1, 0.0, np.inf, download=False) if "download" in ctx.knn.__code__.co_varnames else ctx.knn(None, 1, 0.0, np.inf)
    ctx.timing_enable(True)
    for name, chain in (("TrimmedDist", [(capi.FILTER_TRIMMEDDIST, 0.75)]), ("VarTrimmedDist", [(capi.FILTER_VARTRIMMEDDIST, 2.35)])):
        ctx.weights(chain, download=False); ctx.timing_collect()
        for _ in range(3):
            ctx.weights(chain, download=False)
        ms, n = ctx.timing_collect()["select"]
        print("%s: %.3f ms per evaluation at 1 M matches" % (name, ms / n), "ratio" if name.startswith("Var") else "", ctx.var_trimmed_ratio() if name.startswith("Var") else "")
