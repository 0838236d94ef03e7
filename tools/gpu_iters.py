"""Developer probe: per-scenario PDHG iteration counts of a batch (to study the tail).  python tools/gpu_iters.py case118 128 [k=v ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.build()
import bench
from activesetmethods_b200.sublp import SubLp
case, S = sys.argv[1], int(sys.argv[2])
opts = {}
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    opts[k] = float(v) if ("." in v or "e" in v) else int(v)
net = bench.network(case)
mdl, d = bench.linearise(net, [1 + s for s in range(S)])
lp = SubLp(mdl.n, mdl.m, mdl.j_str, d["xL"], d["xU"], d["gL"], d["gU"], batch=S, **{"eps_rel": 1e-6, **opts})
lp.sub_optimize(d["x"], d["f"], d["df"], d["E"], d["dE"], 1000.0, False)
its = np.array([i["iterations"] for i in lp.last_info]); st = np.array([i["status"] for i in lp.last_info])
order = np.argsort(-its)
print(case, S, opts, "mean", its.mean(), "median", np.median(its), "max", its.max(), "statuses", np.bincount(st))
print("worst:", [(int(o) + 1, int(its[o])) for o in order[:10]])
print("quantiles 50/75/90/95/99:", [int(np.quantile(its, q)) for q in (0.5, 0.75, 0.9, 0.95, 0.99)])
ms, _ = lp.last_solve_timing()
print("loop ms", ms)
