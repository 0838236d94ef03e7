"""Developer probe (not a test, not the bench): per-iteration device time of the PDHG loop for (case, batch)
pairs, iteration-capped so it finishes quickly.  Usage: python tools/gpu_probe.py case1354pegase 1,64,256,1024 [iters]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

g.build()
from activesetmethods_b200.examples import acopf  # noqa: E402
from activesetmethods_b200.sublp import SubLp  # noqa: E402

name = sys.argv[1]
batches = [int(b) for b in sys.argv[2].split(",")]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
opts = {}
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    opts[k] = float(v) if ("." in v or "e" in v) else int(v)
net = acopf.case9() if name == "case9" else acopf.synthetic_network(*acopf.PEGASE_SHAPES[name])
base = acopf.AcopfModel(net)
n, m, nnz = base.n, base.m, len(base.j_str)
b_iter = 24 * nnz + 76 * n + 60 * m + 8
for B in batches:
    t0 = time.time()
    xs, fs, dfs, Es, dEs, gL, gU = [], [], [], [], [], [], []
    for s in range(B):
        mdl = base if s == 0 else acopf.AcopfModel(acopf.perturb_loads(net, s))
        x = np.clip(mdl.x0, mdl.x_L, mdl.x_U)
        xs.append(x); fs.append(mdl.eval_f(x)); dfs.append(mdl.eval_grad_f(x, np.zeros(n)))
        Es.append(mdl.eval_g(x, np.zeros(m))); dEs.append(mdl.eval_jac_g(x, "eval", None, None, np.zeros(nnz)))
        gL.append(mdl.g_L); gU.append(mdl.g_U)
    t_host = time.time() - t0
    per = B > 1
    lp = SubLp(n, m, base.j_str, np.tile(base.x_L, (B, 1)) if per else base.x_L, np.tile(base.x_U, (B, 1)) if per else base.x_U,
               np.array(gL) if per else gL[0], np.array(gU) if per else gU[0], batch=B, max_iter=iters, eps_rel=1e-30, **opts)
    for rep in range(2):
        t0 = time.time()
        lp.sub_optimize(np.array(xs), np.array(fs), np.array(dfs), np.array(Es), np.array(dEs), 1000.0, False)
        t_call = time.time() - t0
    ms, its = lp.last_solve_timing()
    us = ms * 1e3 / max(its, 1)
    print(f"{name} B={B}: n {n} m {m} nnz {nnz} | host eval {t_host:.1f}s | call {t_call*1e3:.1f} ms loop {ms:.1f} ms its {its} "
          f"-> {us:.2f} us/it, {us / B:.3f} us/scenario-it, algorithmic {b_iter * B / (us * 1e-6) / 1e9:.0f} GB/s", flush=True)
    lp.close()
