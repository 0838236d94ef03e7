"""Developer probe for the barrier engine (engine 4): status / objective / Newton steps / device time of the first
SLP sub-LP of a case, against HiGHS (oracle) where that is quick.
Usage: python tools/gpu_ipm_probe.py case[,case...] batch[,batch...] [key=value ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
from activesetmethods_b200.examples import acopf, small_nlps  # noqa: E402
from activesetmethods_b200.sublp import SubLp  # noqa: E402
from oracle import slp_oracle as so  # noqa: E402

names = sys.argv[1].split(",")
batches = [int(b) for b in sys.argv[2].split(",")]
opts = {}
oracle_max = 3000
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    if k == "oracle_max":
        oracle_max = int(v)
        continue
    opts[k] = float(v) if ("." in v or "e" in v) else int(v)
opts.setdefault("engine", 4)
for name in names:
    if name == "toy":
        models = lambda s: small_nlps.ToyNlp()
    elif name == "case9":
        net = acopf.case9()
        models = lambda s: acopf.AcopfModel(net if s == 0 else acopf.perturb_loads(net, s))
    else:
        net = acopf.synthetic_network(*acopf.PEGASE_SHAPES[name])
        models = lambda s: acopf.AcopfModel(net if s == 0 else acopf.perturb_loads(net, s))
    base = models(0)
    n, m, nnz = base.n, base.m, len(base.j_str)
    for B in batches:
        xs, fs, dfs, Es, dEs, gL, gU = [], [], [], [], [], [], []
        for s in range(B):
            mdl = base if s == 0 else models(s)
            x = np.clip(mdl.x0, mdl.x_L, mdl.x_U)
            xs.append(x); fs.append(mdl.eval_f(x)); dfs.append(mdl.eval_grad_f(x, np.zeros(n)))
            Es.append(mdl.eval_g(x, np.zeros(m))); dEs.append(mdl.eval_jac_g(x, "eval", None, None, np.zeros(nnz)))
            gL.append(mdl.g_L); gU.append(mdl.g_U)
        per = B > 1
        t0 = time.time()
        lp = SubLp(n, m, base.j_str, np.tile(base.x_L, (B, 1)) if per else base.x_L,
                   np.tile(base.x_U, (B, 1)) if per else base.x_U, np.array(gL) if per else gL[0],
                   np.array(gU) if per else gU[0], batch=B, **opts)
        t_create = time.time() - t0
        for rep in range(2):
            t0 = time.time()
            out = lp.sub_optimize(np.array(xs), np.array(fs), np.array(dfs), np.array(Es), np.array(dEs), 1000.0, False)
            t_call = time.time() - t0
            if rep == 0:
                t_first = t_call
                if "verbose" in opts:
                    lp.params.verbose = 0
        ms, its = lp.last_solve_timing()
        info = lp.last_info
        st = [i["status"] for i in info]
        it = [i["iterations"] for i in info]
        print(f"{name} B={B}: n {n} m {m} nnz {nnz} | create {t_create:.2f}s first call {t_first:.2f}s call {t_call*1e3:.1f} ms "
              f"loop {ms:.1f} ms | status {sorted(set(st))} newton {min(it)}..{max(it)} obj0 {info[0]['objective']:.10f} "
              f"pres {max(i['primal_residual'] for i in info):.2e} dres {max(i['dual_residual'] for i in info):.2e} "
              f"gap {max(i['gap'] for i in info):.2e} | launches {lp.launch_count()}", flush=True)
        if n <= oracle_max:
            pat = so.JacobianPattern(base.m, base.n, base.j_str)
            worst = 0.0
            for s in range(min(B, 4)):
                mdl = base if s == 0 else models(s)
                ref = so.SubLp(pat, mdl.g_L, mdl.g_U, mdl.x_L, mdl.x_U)
                rs = ref.solve(pat.assemble(dEs[s]), dfs[s], fs[s], Es[s], xs[s], 1000.0, False)[5]
                if rs == 0 and st[s] == 0:
                    worst = max(worst, abs(info[s]["objective"] - ref.last_objective) / max(1.0, abs(ref.last_objective)))
                print(f"   scenario {s}: oracle status {rs} objective {ref.last_objective}  gpu status {st[s]} objective {info[s]['objective']:.10f}")
            print(f"   worst relative objective difference {worst:.2e}")
        lp.close()
