"""Generator of tests/golden/sublp_case118_tinyrow.npz (needs a GPU: it records the second sub-LP of the *GPU* SLP
run on the synthetic case118, whose iterate leaves two thermal-limit rows with gradient ~3e-11 — the input that
exposed the equilibration blow-up fixed by kTinyRel in csrc/lp_solver.cuh).  The expected status / objective are
the oracle's simplex solve of that recorded linearisation.

    gpurun -- python tests/golden/make_tinyrow.py      # writes gpurun_out/lp_dump.npz; then, on any box:
    python tests/golden/make_tinyrow.py --finish
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import problem  # noqa: E402

DUMP = os.path.join(ROOT, "gpurun_out", "lp_dump.npz")

if "--finish" not in sys.argv:
    import __graft_entry__ as g
    g.build()
    from activesetmethods_b200.slp import Model, Parameters, SlpLS
    pr = problem("case118")
    mdl = Model.from_problem(pr, Parameters(algorithm="Line Search", max_iter=3,
                                            lp_options=dict(eps_rel=1e-6, max_iter=200000, warm_start=0)))
    slp = SlpLS(mdl)
    rec = []
    slp.record = lambda s, d: rec.append(d)
    slp.run()
    os.makedirs(os.path.dirname(DUMP), exist_ok=True)
    np.savez(DUMP, **{f"{k}_{i}": np.asarray(v) for i, d in enumerate(rec) for k, v in d.items()})
else:
    from oracle import slp_oracle as so
    z = np.load(DUMP)
    pr = problem("case118")
    pat = so.JacobianPattern(pr.m, pr.n, pr.j_str)
    ref = so.SubLp(pat, pr.g_L, pr.g_U, pr.x_L, pr.x_U)
    out = ref.solve(pat.assemble(z["dE_1"]), z["df_1"], float(z["f_1"]), z["E_1"], z["x_1"], 1000.0, False)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sublp_case118_tinyrow.npz"), x=z["x_1"], f=z["f_1"],
                        df=z["df_1"], E=z["E_1"], dE=z["dE_1"], delta=1000.0, fr=False, status=out[5],
                        objective=ref.last_objective)
