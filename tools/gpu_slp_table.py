"""Developer tool: whole SLP solves, GPU engine vs the oracle (HiGHS dual simplex underneath), one line per
(problem, algorithm, tolerance set).
python tools/gpu_slp_table.py toy,hs071,case9 LS,TR loose,tight [oracle=0] [max_iter=N] [k=v lp options ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
g.build()
from helpers import problem
from oracle import slp_oracle as so
from activesetmethods_b200.slp import Model, Parameters, SlpLS, SlpTR

names = sys.argv[1].split(",")
algs = sys.argv[2].split(",")
tols = sys.argv[3].split(",")
opts = {}
run_oracle = 1
max_iter = 300
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    if k == "oracle":
        run_oracle = int(v)
    elif k == "max_iter":
        max_iter = int(v)
    else:
        opts[k] = float(v) if ("." in v or "e" in v) else int(v)
TOL = {"loose": dict(), "tight": dict(tol_residual=1e-8, tol_infeas=1e-8)}
for name in names:
    for a in algs:
        alg = {"LS": "Line Search", "TR": "Trust Region"}[a]
        for tl in tols:
            line = f"{name:16s} {a} {tl:5s}"
            if run_oracle:
                t0 = time.time()
                ref = (so.SlpLS if a == "LS" else so.SlpTR)(problem(name), so.Parameters(algorithm=alg, max_iter=max_iter, **TOL[tl]))
                ref.run()
                line += f" | oracle ret {ref.ret:3d} it {ref.iter:4d} obj {ref.obj_val:.9f} viol {ref.prim_infeas:.2e} {time.time()-t0:6.1f}s"
            t0 = time.time()
            mdl = Model.from_problem(problem(name), Parameters(algorithm=alg, max_iter=max_iter, lp_options=dict(opts), **TOL[tl]))
            slp = (SlpLS if a == "LS" else SlpTR)(mdl)
            recs = []
            slp.record = lambda s_, d: recs.append(d)
            slp.run()
            for k, e in enumerate(slp.lp_log):
                if e[0] > 1 and k < len(recs):
                    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                    np.savez(os.path.join(ROOT, "gpurun_out", f"badlp_{name}_{a}_{tl}_{k}.npz"), status=e[0], **recs[k])
                    print(f"   sub-LP {k}: status {e[0]} after {e[3]} Newton steps, delta {recs[k]['delta']}, restoration {recs[k]['fr']}")
            dt = time.time() - t0
            st = [e[0] for e in slp.lp_log]
            its = [e[3] for e in slp.lp_log]
            line += (f" | gpu ret {slp.ret:3d} it {slp.iter:4d} obj {slp.obj_val:.9f} viol {slp.prim_infeas:.2e} {dt:6.1f}s "
                     f"LPs {len(st)} (infeasible {st.count(1)}, other {sum(1 for s in st if s > 1)}) newton {sum(its)}")
            if run_oracle and ref.ret == slp.ret:
                line += f" | rel obj diff {abs(ref.obj_val - slp.obj_val) / max(1.0, abs(ref.obj_val)):.2e}"
            print(line, flush=True)
