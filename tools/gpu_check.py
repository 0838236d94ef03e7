"""Developer diagnostic (not a test): solve recorded sub-LPs of a case on the GPU and print, per LP, status /
objective / iterations against the oracle.  Usage: python tools/gpu_check.py case9 [LS|TR] [limit] [key=value ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as g  # noqa: E402

g.build()
from helpers import problem, record_sublps  # noqa: E402
from oracle import slp_oracle as so  # noqa: E402
from activesetmethods_b200.sublp import SubLp  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "case9"
alg = {"LS": "Line Search", "TR": "Trust Region"}[sys.argv[2] if len(sys.argv) > 2 else "LS"]
limit = int(sys.argv[3]) if len(sys.argv) > 3 else 6
opts = {}
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    opts[k] = float(v) if ("." in v or "e" in v) else int(v)
pr = problem(name)
t = time.time()
NOREF = bool(int(os.environ.get("NOREF", "0")))
if limit == 0:   # first linearisation only, no oracle SLP run (big cases)
    x = np.clip(pr.x0, pr.x_L, pr.x_U)
    lps = [dict(x=x, f=pr.eval_f(x), df=pr.eval_grad_f(x, np.zeros(pr.n)), E=pr.eval_g(x, np.zeros(pr.m)),
                dE=pr.eval_jac_g(x, "eval", None, None, np.zeros(len(pr.j_str))),
                delta=1000.0 if alg == "Line Search" else 0.4, fr=False)]

    class _S:
        ret = iter = -1
        obj_val = float("nan")
    slp = _S()
else:
    slp, lps = record_sublps(pr, alg, max_iter=40, limit=limit)
print(f"{name} {alg}: n {pr.n} m {pr.m} nnz {len(pr.j_str)}; oracle SLP ret {slp.ret} iters {slp.iter} obj {slp.obj_val:.6f} "
      f"in {time.time() - t:.2f}s", flush=True)
pat = so.JacobianPattern(pr.m, pr.n, pr.j_str)
lp = SubLp(pr.n, pr.m, pr.j_str, pr.x_L, pr.x_U, pr.g_L, pr.g_U, **opts)
for k, d in enumerate(lps):
    ref = so.SubLp(pat, pr.g_L, pr.g_U, pr.x_L, pr.x_U)
    t0 = time.time()
    ro = (None,) * 5 + (-1,) if NOREF else ref.solve(pat.assemble(d["dE"]), d["df"], d["f"], d["E"], d["x"], d["delta"], d["fr"])
    t_ref = time.time() - t0
    t0 = time.time()
    out = lp.sub_optimize(d["x"], d["f"], d["df"], d["E"], d["dE"], d["delta"], d["fr"])
    t_gpu = time.time() - t0
    i = lp.last_info[0]
    ms, its = lp.last_solve_timing()
    rel = abs(i["objective"] - ref.last_objective) / max(1.0, abs(ref.last_objective)) if ro[5] == 0 and out[5] == 0 else float("nan")
    print(f"LP {k} fr={int(d['fr'])} delta={d['delta']:.3g}: gpu status {out[5]} obj {i['objective']:.9e} it {i['iterations']} "
          f"restarts {i['restarts']} pres {i['primal_residual']:.1e} dres {i['dual_residual']:.1e} gap {i['gap']:.1e} | "
          f"oracle status {ro[5]} obj {ref.last_objective if ref.last_objective is not None else float('nan'):.9e} | rel {rel:.2e} | "
          f"gpu {t_gpu * 1e3:.1f} ms (loop {ms:.1f} ms, {ms * 1e3 / max(its, 1):.2f} us/it) highs {t_ref * 1e3:.1f} ms", flush=True)
