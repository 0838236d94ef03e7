"""Device-side ACOPF evaluator against the numpy evaluator that feeds every other test
(activesetmethods_b200/examples/acopf.py): f, grad f, g and the Jacobian values in j_str order at random points, for
a network with a dc line, taps / shifts and shunts (case3), case9, and a batch of the synthetic case118; then the
trial merit phi(x + alpha p) of the line search against its host computation.  Tolerance 1e-12 relative: the device
uses sincos / FMA where numpy uses separate sin, cos and multiplies."""
import numpy as np
import pytest

from activesetmethods_b200.examples import acopf
from helpers import problem
from test_oracle_pins import case3_network

pytestmark = pytest.mark.gpu


def _close(a, b, name):
    a, b = np.asarray(a), np.asarray(b)
    err = np.max(np.abs(a - b) / (1.0 + np.abs(b)), initial=0.0)
    assert err <= 1e-12, (name, err)


@pytest.mark.parametrize("name,B", [("case3", 1), ("case9", 1), ("case118", 33)])
def test_device_evaluator_matches_host(gpu, name, B):
    from activesetmethods_b200.sublp import SubLp
    mdl = acopf.AcopfModel(case3_network()) if name == "case3" else problem(name)
    rng = np.random.default_rng(3)
    lo = np.where(np.isfinite(mdl.x_L), mdl.x_L, -0.5)
    hi = np.where(np.isfinite(mdl.x_U), mdl.x_U, 0.5)
    x = rng.uniform(lo, hi, (B, mdl.n))
    lp = SubLp(mdl.n, mdl.m, mdl.j_str, mdl.x_L, mdl.x_U, mdl.g_L, mdl.g_U, batch=B, eps_rel=1e-7)
    lp._squeeze = False
    lp.attach_acopf(mdl)
    lp.eval_acopf(x, 1000.0, False)
    f, df, E, dE = lp.get_eval()
    for s in range(B):
        _close(f[s], mdl.eval_f(x[s]), "f")
        _close(df[s], mdl.eval_grad_f(x[s], np.zeros(mdl.n)), "df")
        _close(E[s], mdl.eval_g(x[s], np.zeros(mdl.m)), "E")
        _close(dE[s], mdl.eval_jac_g(x[s], "eval", None, None, np.zeros(mdl.nnz)), "dE")
    # the evaluation feeds the same sub-LP as a host hand-over of the same numbers
    x0 = np.tile(np.clip(mdl.x0, mdl.x_L, mdl.x_U), (B, 1))
    lp.eval_acopf(x0, 1000.0, False)
    p, lam, mu_u, mu_l, slack, st = lp.solve_extract()
    obj_dev = np.array([i["objective"] for i in lp.last_info])
    f0, df0, E0, dE0 = lp.get_eval()
    out = lp.sub_optimize(x0, f0, df0, E0, dE0, 1000.0, False)
    obj_host = np.array([i["objective"] for i in lp.last_info])
    assert np.array_equal(np.atleast_1d(st), np.atleast_1d(out[5]))
    ok = np.atleast_1d(st) == 0
    assert np.all(np.abs(obj_dev - obj_host)[ok] <= 1e-6 * np.maximum(1.0, np.abs(obj_host[ok])))
    # trial merit: phi(x + alpha p) = f(x + alpha p) + sum nu * violation(g(x + alpha p))
    if ok.all():
        lp.eval_acopf(x0, 1000.0, False)
        p, lam, mu_u, mu_l, slack, st = lp.solve_extract()
        nu = np.abs(np.atleast_2d(lam))
        alpha = rng.uniform(0.05, 1.0, B)
        phi = np.atleast_1d(lp.acopf_trial(alpha, nu))
        for s in range(B):
            xt = x0[s] + alpha[s] * np.atleast_2d(p)[s]
            Et = mdl.eval_g(xt, np.zeros(mdl.m))
            ref = mdl.eval_f(xt) + float(np.sum(nu[s] * np.maximum(0.0, np.maximum(Et - mdl.g_U, mdl.g_L - Et))))
            assert abs(phi[s] - ref) <= 1e-10 * max(1.0, abs(ref)), (s, phi[s], ref)
    lp.close()
