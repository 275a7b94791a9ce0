"""kNN kernel time versus number of queries against a fixed 1 M-point reference (tail / latency diagnosis)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libpointmatcher_b200 import capi, synth

rd, rf, T_gt = synth.scan_pair(1_000_000)
Tg = T_gt.astype(np.float32)
rng = np.random.default_rng(0)
with capi.Context(0) as ctx:
    ctx.set_reference(rf)
    ctx.timing_enable(True)
    for nq in (31250, 125000, 250000, 500000, 1000000):
        for name, sub in (("prefix", rd[:nq]), ("random", rd[np.sort(rng.choice(len(rd), nq, replace=False))])):
            ctx.set_reading(np.ascontiguousarray(sub))
            for T, tn in ((None, "identity"), (Tg, "aligned")):
                for rep in range(3):
                    ids, d, visits = ctx.knn(T, 1, download=False)
                    ms = ctx.timing_collect()["knn"][0]
                print("nq %8d %-7s %-8s knn %.3f ms  (%.1f Mq/s) visits/query %.1f" % (nq, name, tn, ms, nq / ms / 1e3, visits / nq))
