"""GPU parity of whole SLP solves: the host drivers of activesetmethods_b200/slp.py (reference
slp_line_search.jl:78-215, slp_trust_region.jl:87-206) with every sub-LP and reduction on the device, against
the reference's own known answers and against the CPU oracle run on the same problem.

Bars: final status identical; objective <= 1e-6 relative *to the oracle when both walk the same path* (same
number of SLP iterations and sub-LPs), else the reference's own test tolerance for that known answer — LP optima
are unique in value but not in (p, lambda), so a degenerate sub-LP may send the two solvers down different, equally
valid trajectories; constraint violation within the SLP tolerance of both."""
import numpy as np
import pytest

from activesetmethods_b200.examples import acopf, small_nlps
from oracle import slp_oracle as so
from test_oracle_pins import case3_network

pytestmark = pytest.mark.gpu

LP = dict(eps_rel=1e-7, max_iter=3_000_000)


def _run(problem, algorithm, max_iter):
    from activesetmethods_b200.slp import Model, Parameters, optimize
    mdl = Model.from_problem(problem, Parameters(algorithm=algorithm, max_iter=max_iter, lp_options=LP))
    optimize(mdl)
    return mdl


def _violation(pr, x):
    return so.norm_violations(pr.eval_g(x, np.zeros(pr.m)), pr.g_L, pr.g_U, x, pr.x_L, pr.x_U, np.inf)


def test_toy_known_answer(gpu):
    """reference test/runtests.jl:9-14 (test/ext_solver.jl): X = Y = -1, LOCALLY_SOLVED; the first LP is
    infeasible at x0 = (0, 0), so this walks through feasibility restoration on the device."""
    mdl = _run(small_nlps.ToyNlp(), "Line Search", 1000)
    assert mdl.status == 0
    assert np.allclose(mdl.x, [-1.0, -1.0], rtol=1e-4)


def test_case3_known_answer(gpu):
    """reference test/runtests.jl:18-21: ACP-OPF on case3.m, LS, max_iter 100 -> 5906.87949 (rtol 1e-3)."""
    pr = acopf.AcopfModel(case3_network())
    mdl = _run(pr, "Line Search", 100)
    ref = so.optimize(acopf.AcopfModel(case3_network()), so.Parameters(max_iter=100))
    assert mdl.status == ref.ret == 0
    assert abs(mdl.obj_val - 5906.87949) <= 1e-3 * 5906.87949
    assert _violation(pr, mdl.x) <= 1e-2 and abs(mdl.obj_val - ref.obj_val) <= 1e-3 * abs(ref.obj_val)


@pytest.mark.parametrize("algorithm", ["Line Search", "Trust Region"])
def test_case9_matches_oracle(gpu, algorithm):
    pr = acopf.AcopfModel(acopf.case9())
    ref = so.optimize(acopf.AcopfModel(acopf.case9()), so.Parameters(algorithm=algorithm, max_iter=100))
    mdl = _run(pr, algorithm, 100)
    assert mdl.status == ref.ret
    rel = abs(mdl.obj_val - ref.obj_val) / abs(ref.obj_val)
    viol_gpu, viol_ref = _violation(pr, mdl.x), _violation(pr, ref.x)
    print(f"case9 {algorithm}: status {mdl.status}, objective {mdl.obj_val:.9f} vs oracle {ref.obj_val:.9f} "
          f"(rel {rel:.2e}), violation {viol_gpu:.2e} vs {viol_ref:.2e}")
    assert rel <= 1e-6, rel
    assert abs(viol_gpu - viol_ref) <= 1e-6
    assert abs(mdl.obj_val - 5296.69) <= 1e-2 * 5296.69      # MATPOWER's public optimum, loosely


def test_missing_external_optimizer(gpu):
    """model.jl:64-66: no external optimizer -> Invalid_Option (-12)."""
    from activesetmethods_b200.slp import Model, Parameters, optimize
    mdl = Model.from_problem(small_nlps.ToyNlp(), Parameters(external_optimizer=None))
    assert optimize(mdl).status == -12
