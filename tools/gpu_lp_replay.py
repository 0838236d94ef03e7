"""Developer tool: replay one recorded sub-LP (npz written by tools/gpu_slp_table.py) on the GPU with verbose output.
python tools/gpu_lp_replay.py file.npz problem [k=v lp options ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
g.build()
from helpers import problem
from activesetmethods_b200.sublp import SubLp
d = np.load(sys.argv[1])
pr = problem(sys.argv[2])
opts = dict(verbose=1)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    opts[k] = float(v) if ("." in v or "e" in v) else int(v)
lp = SubLp(pr.n, pr.m, pr.j_str, pr.x_L, pr.x_U, pr.g_L, pr.g_U, **opts)
out = lp.sub_optimize(d["x"], float(d["f"]), d["df"], d["E"], d["dE"], float(d["delta"]), bool(d["fr"]))
print("status", out[5], lp.last_info[0])
