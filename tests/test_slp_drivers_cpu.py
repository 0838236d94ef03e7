"""Host drivers of activesetmethods_b200/slp.py (reference slp_line_search.jl:78-215, slp_trust_region.jl:87-206)
on the CPU: with the oracle's LP solver plugged in through the ``external_optimizer`` hook they must walk exactly the
trajectory of the oracle's own restatement of the reference drivers — same sub-LP sequence, statuses, final status,
iteration count and objective.  Covers feasibility restoration (toy), both algorithms, and the -12 status."""
import numpy as np
import pytest

from activesetmethods_b200.examples import acopf, small_nlps
from activesetmethods_b200.slp import Model, Parameters, SlpLS, SlpLSBatch, SlpTR, optimize, STATUS
from cpu_engine import OracleBatchEngine, OracleEngine
from oracle import slp_oracle as so
from test_oracle_pins import case3_network

CASES = {"toy": small_nlps.ToyNlp, "hs071": small_nlps.Hs071, "case3": lambda: acopf.AcopfModel(case3_network()),
         "case9": lambda: acopf.AcopfModel(acopf.case9())}


@pytest.mark.parametrize("name,algorithm,max_iter", [("toy", "Line Search", 1000), ("case3", "Line Search", 100),
                                                     ("case9", "Line Search", 100), ("case9", "Trust Region", 100),
                                                     ("hs071", "Line Search", 60), ("toy", "Trust Region", 60)])
def test_driver_matches_oracle_trajectory(name, algorithm, max_iter):
    ref_cls = so.SlpLS if algorithm == "Line Search" else so.SlpTR
    ref = ref_cls(CASES[name](), so.Parameters(algorithm=algorithm, max_iter=max_iter))
    ref.run()
    cls = SlpLS if algorithm == "Line Search" else SlpTR
    mdl = Model.from_problem(CASES[name](), Parameters(algorithm=algorithm, max_iter=max_iter,
                                                       external_optimizer=OracleEngine))
    slp = cls(mdl).run()
    assert slp.ret == ref.ret and slp.iter == ref.iter
    assert [(e[0], e[2]) for e in slp.lp_log] == [(e[0], e[2]) for e in ref.lp_log]
    for a, b in zip(slp.lp_log, ref.lp_log):
        if a[0] == 0:
            assert abs(a[1] - b[1]) <= 1e-9 * max(1.0, abs(b[1]))
    assert np.allclose(slp.x, ref.x, rtol=0, atol=1e-10)
    assert abs(mdl.obj_val - ref.obj_val) <= 1e-10 * max(1.0, abs(ref.obj_val))
    assert mdl.status == ref.ret and mdl.status in STATUS


def test_invalid_option_without_optimizer():
    mdl = Model.from_problem(small_nlps.ToyNlp(), Parameters(external_optimizer=None))
    assert optimize(mdl).status == -12 and STATUS[-12] == "Invalid_Option"


def test_trust_region_collapse_terminates():
    """tests/golden/case9_tr_collapsed.npz: a case9 iterate (recorded on a GPU trust-region run) that is feasible to
    tolerance while the trust region has shrunk to 2e-7, so the linearisation is infeasible inside the box.  The
    reference alternates normal / restoration LPs forever from here (slp_trust_region.jl:163-170 `continue`s past its
    max_iter test); both drivers must stop at max_iter with a LOCALLY_SOLVED-class status instead."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "case9_tr_collapsed.npz"))
    pr = acopf.AcopfModel(acopf.case9())
    mdl = Model.from_problem(pr, Parameters(algorithm="Trust Region", max_iter=25, external_optimizer=OracleEngine))
    slp = SlpTR(mdl)
    slp.x = g["x"].copy()
    slp.delta = float(g["delta"])
    slp.run()
    assert slp.ret == 6 and slp.iter == 25 and len(slp.lp_log) <= 2 * 25
    ref = so.SlpTR(acopf.AcopfModel(acopf.case9()), so.Parameters(algorithm="Trust Region", max_iter=25))
    ref.x = g["x"].copy()
    ref.delta = float(g["delta"])
    ref.clip_start = lambda: None
    ref.run()
    assert ref.ret == 6 and ref.iter == 25


def test_batched_line_search_equals_single_runs():
    """SlpLSBatch (lock-step over scenarios, one batched hot-path call per round) must give every scenario exactly what
    its own SlpLS run gives: same status, iteration count and iterate, for load scenarios of case9 and for a batch
    that mixes scenarios entering feasibility restoration (toy, started at different points)."""
    net = acopf.case9()
    probs = [acopf.AcopfModel(acopf.perturb_loads(net, s + 1)) for s in range(4)]
    batch = SlpLSBatch(probs, Parameters(max_iter=100, external_optimizer=OracleBatchEngine)).run()
    for s in range(4):
        one = SlpLS(Model.from_problem(acopf.AcopfModel(acopf.perturb_loads(net, s + 1)),
                                       Parameters(max_iter=100, external_optimizer=OracleEngine))).run()
        assert batch.ret[s] == one.ret and batch.iter[s] == one.iter, (s, batch.ret[s], one.ret)
        assert np.allclose(batch.x[s], one.x, rtol=0, atol=1e-9)
        assert abs(batch.obj_val[s] - one.obj_val) <= 1e-9 * abs(one.obj_val)
    toys = [small_nlps.ToyNlp() for _ in range(3)]
    toys[1].x0 = np.array([0.5, -0.5])
    toys[2].x0 = np.array([-2.0, -0.4])
    tb = SlpLSBatch(toys, Parameters(max_iter=200, external_optimizer=OracleBatchEngine)).run()
    for s in range(3):
        pr = small_nlps.ToyNlp()
        pr.x0 = toys[s].x0.copy()
        one = SlpLS(Model.from_problem(pr, Parameters(max_iter=200, external_optimizer=OracleEngine))).run()
        assert tb.ret[s] == one.ret and tb.iter[s] == one.iter, (s, tb.ret[s], one.ret, tb.iter[s], one.iter)
        assert np.allclose(tb.x[s], one.x, rtol=0, atol=1e-9)
