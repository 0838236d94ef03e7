"""Developer diagnostic: GPU SLP vs oracle SLP, iteration by iteration.  python tools/gpu_slp_compare.py case9 LS [eps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
g.build()
from helpers import problem
from oracle import slp_oracle as so
from activesetmethods_b200.slp import Model, Parameters, SlpLS, SlpTR

name = sys.argv[1]; alg = {"LS": "Line Search", "TR": "Trust Region"}[sys.argv[2]]
eps = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-7
pr = problem(name)
ref_log = []
cls = so.SlpLS if alg == "Line Search" else so.SlpTR
ref = cls(problem(name), so.Parameters(algorithm=alg, max_iter=100))
ref.record = lambda s, d: ref_log.append(d)
ref.run()
mdl = Model.from_problem(pr, Parameters(algorithm=alg, max_iter=100, lp_options=dict(eps_rel=eps, max_iter=3000000)))
slp = (SlpLS if alg == "Line Search" else SlpTR)(mdl)
gpu_log = []
slp.record = lambda s, d: gpu_log.append(d)
slp.run()
print(f"oracle: ret {ref.ret} iters {ref.iter} obj {ref.obj_val:.9f} | gpu: ret {slp.ret} iters {slp.iter} obj {slp.obj_val:.9f}")
for k in range(max(len(ref_log), len(gpu_log))):
    a = ref_log[k] if k < len(ref_log) else None
    b = gpu_log[k] if k < len(gpu_log) else None
    dx = np.max(np.abs(a["x"] - b["x"])) if a is not None and b is not None else float("nan")
    la = ref.lp_log[k] if k < len(ref.lp_log) else None
    lb = slp.lp_log[k] if k < len(slp.lp_log) else None
    print(k, f"|x_ref - x_gpu|inf {dx:.2e}", "ref", None if la is None else (la[0], la[1], la[2]), "gpu", lb)
