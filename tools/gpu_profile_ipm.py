"""Profiling driver (run under ncu): ONE short barrier-engine solve of a case1354 scenario batch.
python tools/gpu_profile_ipm.py [batch] [newton_steps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from activesetmethods_b200.sublp import SubLp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
net = bench.network("case1354pegase")
mdl, d1 = bench.linearise(net, [1])
# the profile only needs representative values: every scenario gets scenario 1's linearisation
lp = SubLp(mdl.n, mdl.m, mdl.j_str, mdl.x_L, mdl.x_U, mdl.g_L, mdl.g_U, batch=B, engine=4, ipm_max_iter=steps)
rep = lambda a: np.repeat(a, B, axis=0)
out = lp.sub_optimize(rep(d1["x"]), rep(d1["f"]), rep(d1["df"]), rep(d1["E"]), rep(d1["dE"]), 1000.0, False)
print("status", sorted(set(int(s) for s in np.atleast_1d(out[5]))), lp.ipm_info(), "launches", lp.launch_count())
lp.close()
