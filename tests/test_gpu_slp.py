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


def test_case9_trust_region(gpu):
    """Trust region on case9.  The first sub-LP is infeasible, and the restoration LPs that follow (min sum of
    slacks) have a whole face of minimisers: the oracle's simplex vertex and the PDHG point differ after the second
    LP (|dx| ~ 0.3), so the two runs are different, equally valid SLP trajectories (the reference authors call this
    driver "numerically instable", test/MOI_wrapper.jl:90,106).  Required: every sub-LP solved to optimality or
    proven infeasible exactly where the oracle's is on the common prefix, a LOCALLY_SOLVED final status (0 or 6, both
    map to it: MOI_wrapper.jl:1158-1200), the public optimum 5296.69 to the reference's suite tolerance, feasibility."""
    from activesetmethods_b200.slp import Model, Parameters, SlpTR
    pr = acopf.AcopfModel(acopf.case9())
    ref = so.SlpTR(acopf.AcopfModel(acopf.case9()), so.Parameters(algorithm="Trust Region", max_iter=100))
    ref.run()
    slp = SlpTR(Model.from_problem(pr, Parameters(algorithm="Trust Region", max_iter=100, lp_options=LP))).run()
    assert ref.ret == 0 and slp.ret in (0, 6)
    # common prefix: same inputs -> same LP status and objective (LP 0 infeasible, LP 1 restoration optimum)
    assert slp.lp_log[0][0] == ref.lp_log[0][0] == so.INFEASIBLE
    assert slp.lp_log[1][0] == ref.lp_log[1][0] == 0 and slp.lp_log[1][2] and ref.lp_log[1][2]
    assert abs(slp.lp_log[1][1] - ref.lp_log[1][1]) <= 1e-6 * max(1.0, abs(ref.lp_log[1][1]))
    assert all(entry[0] in (0, 1) for entry in slp.lp_log)
    print(f"case9 TR: status {slp.ret}, objective {slp.obj_val:.6f} vs oracle {ref.obj_val:.6f}")
    assert abs(slp.obj_val - 5296.69) <= 1e-2 * 5296.69
    assert _violation(pr, slp.x) <= 1e-2


def test_case9_line_search_status_and_tolerance(gpu):
    """Line search stops at the reference's loose default tolerances (tol_residual = tol_infeas = 1e-2,
    parameters.jl:18-19) well before the optimum, and the first sub-LP already has a non-unique minimiser (the
    reactive injections carry no cost), so a simplex vertex and a PDHG point send the two runs down different
    paths (the oracle itself ends at 5279.70, 3e-3 below the optimum).  What must hold: every sub-LP the GPU run
    solved is optimal, the final status is the oracle's, both end within the reference's suite tolerance (1e-2,
    test/MOI_wrapper.jl:14) of the optimum and feasible to tol_infeas."""
    from activesetmethods_b200.slp import Model, Parameters, SlpLS
    pr = acopf.AcopfModel(acopf.case9())
    ref = so.optimize(acopf.AcopfModel(acopf.case9()), so.Parameters(max_iter=100))
    slp = SlpLS(Model.from_problem(pr, Parameters(max_iter=100, lp_options=LP))).run()
    assert slp.ret == ref.ret == 0
    assert all(entry[0] == 0 for entry in slp.lp_log)
    for obj in (slp.obj_val, ref.obj_val):
        assert abs(obj - 5296.69) <= 1e-2 * 5296.69
    assert _violation(pr, slp.x) <= 1e-2


def test_batched_scenarios_line_search(gpu):
    """Six load scenarios of case9 solved together by the lock-step batched driver (one batched device call per SLP
    round) : every scenario ends LOCALLY_SOLVED, feasible, and within the reference's suite tolerance of the optimum
    the oracle's trust-region run finds for that scenario."""
    from activesetmethods_b200.slp import Parameters, SlpLSBatch
    net = acopf.case9()
    ids = [1, 2, 3, 4, 5, 6]
    probs = [acopf.AcopfModel(acopf.perturb_loads(net, s)) for s in ids]
    batch = SlpLSBatch(probs, Parameters(max_iter=100, lp_options=LP)).run()
    assert batch.rounds <= 120 and batch.lp_iterations > 0
    for k, s in enumerate(ids):
        ref = so.optimize(acopf.AcopfModel(acopf.perturb_loads(net, s)),
                          so.Parameters(algorithm="Trust Region", max_iter=200, tol_residual=1e-4, tol_infeas=1e-4))
        assert batch.ret[k] in (0, 6), (s, batch.ret[k])
        assert _violation(probs[k], batch.x[k]) <= 1e-2
        assert ref.ret in (0, 6)
        assert abs(batch.obj_val[k] - ref.obj_val) <= 1e-2 * abs(ref.obj_val), (s, batch.obj_val[k], ref.obj_val)


def test_batched_scenarios_device_evaluator(gpu):
    """The same batch with the NLP callbacks replaced by the device-side ACOPF evaluator (no host evaluation in the
    loop, backtracking trials evaluated on the device): same statuses and SLP iteration counts as the host-evaluator
    run, objectives equal to 1e-6."""
    from activesetmethods_b200.slp import Parameters, SlpLSBatch
    net = acopf.case9()
    ids = [1, 2, 3, 4, 5, 6]
    host = SlpLSBatch([acopf.AcopfModel(acopf.perturb_loads(net, s)) for s in ids],
                      Parameters(max_iter=100, lp_options=LP)).run()
    dev = SlpLSBatch([acopf.AcopfModel(acopf.perturb_loads(net, s)) for s in ids],
                     Parameters(max_iter=100, lp_options=LP), device_evaluator=True).run()
    assert np.array_equal(dev.ret, host.ret) and np.all(np.isin(dev.ret, (0, 6)))
    assert np.array_equal(dev.iter, host.iter)
    assert np.all(np.abs(dev.obj_val - host.obj_val) <= 1e-6 * np.abs(host.obj_val))


def test_line_search_on_device_evaluator(gpu):
    """SlpLS on case9 with the NLP callbacks replaced by the device-side evaluator: same status, iteration count
    and objective (1e-6) as the host-callback run, and the finish step's E / obj come out right."""
    from activesetmethods_b200.slp import Model, Parameters, SlpLS
    host = SlpLS(Model.from_problem(acopf.AcopfModel(acopf.case9()), Parameters(max_iter=100, lp_options=LP))).run()
    dev = SlpLS(Model.from_problem(acopf.AcopfModel(acopf.case9()),
                                   Parameters(max_iter=100, lp_options=LP, device_evaluator=True))).run()
    assert dev.ret == host.ret == 0 and dev.iter == host.iter
    assert abs(dev.obj_val - host.obj_val) <= 1e-6 * abs(host.obj_val)


def test_missing_external_optimizer(gpu):
    """model.jl:64-66: no external optimizer -> Invalid_Option (-12)."""
    from activesetmethods_b200.slp import Model, Parameters, optimize
    mdl = Model.from_problem(small_nlps.ToyNlp(), Parameters(external_optimizer=None))
    assert optimize(mdl).status == -12
