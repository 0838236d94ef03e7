"""Pins the CPU oracle against the reference's own known answers and the committed fixtures.
CPU only."""
import json
import os

import numpy as np
import pytest

from activesetmethods_b200.examples import acopf, small_nlps
from oracle import slp_oracle as so

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def case3_network():
    d = json.load(open(os.path.join(GOLDEN, "case3_network.json")))
    kw = {k: (np.array(v) if isinstance(v, list) else v) for k, v in d.items()}
    for k in ("bus_id", "gen_bus", "f_bus", "t_bus", "dc_f", "dc_t"):
        kw[k] = np.asarray(kw[k], dtype=np.int64)
    return acopf.Network(**kw)


def test_toy_known_answer():
    """reference test/runtests.jl:9-14: X = Y = -1, LOCALLY_SOLVED."""
    r = so.optimize(small_nlps.ToyNlp(), so.Parameters())
    assert r.ret == 0
    assert np.allclose(r.x, [-1.0, -1.0], rtol=1e-4)


def test_toy_first_lp_infeasible_then_fr_optimum_4():
    """at x0 = (0, 0) the normal LP is INFEASIBLE and the restoration LP optimum is 4.0."""
    pr = small_nlps.ToyNlp()
    slp = so.SlpLS(pr, so.Parameters(max_iter=3))
    slp.run()
    assert slp.lp_log[0][0] == so.INFEASIBLE
    assert slp.lp_log[1][2] is True and abs(slp.lp_log[1][1] - 4.0) < 1e-9


def test_case3_known_answer():
    """reference test/runtests.jl:18-21: ACP-OPF on case3.m, LS, max_iter 100 -> 5906.87949 (rtol 1e-3)."""
    r = so.optimize(acopf.AcopfModel(case3_network()), so.Parameters(max_iter=100))
    assert r.ret == 0
    assert abs(r.obj_val - 5906.87949) <= 1e-3 * 5906.87949


def test_case9_public_optimum_tr():
    """MATPOWER case9 AC-OPF optimum 5296.69 (public value) with the trust-region driver."""
    r = so.optimize(acopf.AcopfModel(acopf.case9()), so.Parameters(algorithm="Trust Region", max_iter=100))
    assert r.ret == 0
    assert abs(r.obj_val - 5296.69) <= 1e-3 * 5296.69


def test_hs071():
    """hs071 (MOIT.nlptest, reference test/MOI_wrapper.jl:109): optimum 17.0140173 at 1e-2 like the suite."""
    r = so.optimize(small_nlps.Hs071(), so.Parameters(max_iter=3000))
    assert r.ret in (0, 6)
    assert abs(r.obj_val - 17.0140173) <= 1e-2 * 17.0140173


@pytest.mark.parametrize("name", ["toy", "case3", "case9", "case9tr"])
def test_golden_sublps_reproduce(name):
    """The committed sub-LP fixtures are what the oracle computes today (guards oracle drift)."""
    g = np.load(os.path.join(GOLDEN, f"sublp_{name}.npz"))
    pr = {"toy": small_nlps.ToyNlp, "case3": lambda: acopf.AcopfModel(case3_network()),
          "case9": lambda: acopf.AcopfModel(acopf.case9()), "case9tr": lambda: acopf.AcopfModel(acopf.case9())}[name]()
    pat = so.JacobianPattern(pr.m, pr.n, pr.j_str)
    for k in range(len(g["f"])):
        sub = so.SubLp(pat, pr.g_L, pr.g_U, pr.x_L, pr.x_U)          # cold start: no basis carried over
        out = sub.solve(pat.assemble(g["dE"][k]), g["df"][k], float(g["f"][k]), g["E"][k], g["x"][k],
                        float(g["delta"][k]), bool(g["fr"][k]))
        assert out[5] == int(g["status"][k])
        if out[5] == so.OPTIMAL:
            assert abs(sub.last_objective - g["objective"][k]) <= 1e-9 * max(1.0, abs(g["objective"][k]))


def test_golden_tinyrow_fixture_reproduces():
    """tests/golden/sublp_case118_tinyrow.npz (recorded on the GPU SLP path) is what the oracle computes today."""
    from helpers import problem
    g = np.load(os.path.join(GOLDEN, "sublp_case118_tinyrow.npz"))
    pr = problem("case118")
    pat = so.JacobianPattern(pr.m, pr.n, pr.j_str)
    sub = so.SubLp(pat, pr.g_L, pr.g_U, pr.x_L, pr.x_U)
    out = sub.solve(pat.assemble(g["dE"]), g["df"], float(g["f"]), g["E"], g["x"], float(g["delta"]), bool(g["fr"]))
    assert out[5] == int(g["status"]) == so.OPTIMAL
    assert abs(sub.last_objective - float(g["objective"])) <= 1e-9 * abs(float(g["objective"]))


def test_assemble_matches_literal_loop():
    """JacobianPattern.assemble == the scalar J[r,c] += v loop of common.jl:15-18, bit for bit, with duplicates."""
    rng = np.random.default_rng(7)
    pr = small_nlps.RandomNlp(n=12, m=9, seed=3) if "seed" in small_nlps.RandomNlp.__init__.__code__.co_varnames \
        else small_nlps.RandomNlp()
    pat = so.JacobianPattern(pr.m, pr.n, pr.j_str)
    assert pat.nnz < len(pr.j_str), "RandomNlp must contain duplicate COO entries"
    for _ in range(5):
        dE = rng.standard_normal(len(pr.j_str)) * 10.0 ** rng.integers(-8, 8, len(pr.j_str))
        ref = so.assemble_reference_loop(pr.m, pr.n, pr.j_str, dE)
        vals = pat.assemble(dE)
        for s in range(pat.nnz):
            assert vals[s] == ref[(int(pat.rows[s]) + 1, int(pat.cols[s]) + 1)]


def test_metrics_small_hand_case():
    """common.jl:35-98 on a hand-checkable case."""
    E = np.array([1.0, -2.0, 0.5]); gl = np.array([0.0, -1.0, 0.5]); gu = np.array([0.5, np.inf, 0.5])
    x = np.array([2.0, -3.0]); xl = np.array([0.0, -1.0]); xu = np.array([1.0, 1.0])
    assert so.norm_violations(E, gl, gu, x, xl, xu, np.inf) == 2.0
    assert so.norm_violations(E, gl, gu, x, xl, xu, 1) == 0.5 + 1.0 + 1.0 + 2.0
    lam = np.array([2.0, 3.0, 7.0])
    # inequality rows 0, 1: min(E-gl, gu-E)*lam = [-0.5*2, -1*3] -> max abs 3 ; denom 1 + sqrt(4 + 9)
    assert abs(so.norm_complementarity(E, gl, gu, lam) - 3.0 / (1.0 + np.sqrt(13.0))) < 1e-15
