"""Parity at BASELINE.json's full sizes through properties that do not need a second solver: the KKT certificate
of every returned sub-LP solution is re-computed on the host from (p, lambda, column duals) with scipy, independent
of the solver's own residuals.  For min c'p, rl <= Jp <= ru, lb <= p <= ub (MOI sign convention consumed at
src/algorithms/common.jl:38):

    primal:  max violation of rows and bounds                                  <= 1e-6 (1 + |b|_inf)
    dual:    |c - J'lambda - z_L - z_U|_2 <= 1e-6 (1 + |c|_2), z_L >= 0 / z_U <= 0 where p sits on its bound
    gap:     |c'p - (rl'lambda+ + ru'lambda- + lb'z_L + ub'z_U)|               <= 1e-6 (1 + |pobj| + |dobj|)

(the termination criteria of the engine, re-derived with independent arithmetic) plus the strict weak-duality lower
bound with every reduced cost priced on the box.  The Jacobian
assembly is compared bit for bit with the restatement of common.jl:12-20.  case1354 is also checked against the
oracle's simplex objective (12 s of HiGHS); case13659 (191 055 rows) only through the certificate."""
import numpy as np
import pytest

from helpers import problem
from oracle import slp_oracle as so

pytestmark = pytest.mark.gpu


def _first_linearisation(pr):
    x = np.clip(pr.x0, pr.x_L, pr.x_U)
    return dict(x=x, f=pr.eval_f(x), df=pr.eval_grad_f(x, np.zeros(pr.n)), E=pr.eval_g(x, np.zeros(pr.m)),
                dE=pr.eval_jac_g(x, "eval", None, None, np.zeros(len(pr.j_str))))


def _certificate(pr, d, delta, lp, p, lam):
    """Host-side KKT certificate from the engine's primal step and row duals (column duals = reduced costs)."""
    pat = so.JacobianPattern(pr.m, pr.n, pr.j_str)
    vals = pat.assemble(d["dE"])
    rp, ci, v = lp.jacobian_csr()
    assert np.array_equal(rp, pat.row_ptr) and np.array_equal(ci, pat.cols)
    assert np.array_equal(v, vals), "device assembly differs from compute_jacobian_matrix"
    J = pat.matrix(vals)
    lb = np.maximum(-delta, pr.x_L - d["x"]); ub = np.minimum(delta, pr.x_U - d["x"])
    rl = np.where(np.isfinite(pr.g_L), pr.g_L - d["E"], -np.inf)
    ru = np.where(np.isfinite(pr.g_U), pr.g_U - d["E"], np.inf)
    Jp = J @ p
    scale = 1.0 + max(np.max(np.abs(rl[np.isfinite(rl)]), initial=0.0), np.max(np.abs(ru[np.isfinite(ru)]), initial=0.0))
    pviol = max(np.max(np.maximum(0.0, rl - Jp)), np.max(np.maximum(0.0, Jp - ru)), np.max(np.maximum(0.0, lb - p)),
                np.max(np.maximum(0.0, p - ub)))
    # row duals must respect the sides that exist
    sign_bad = max(np.max(np.where(~np.isfinite(rl), np.maximum(lam, 0.0), 0.0)),
                   np.max(np.where(~np.isfinite(ru), np.maximum(-lam, 0.0), 0.0)))
    z = d["df"] - J.T @ lam                       # reduced costs
    pobj = d["df"] @ p
    # (i) strict certificate: every reduced cost priced on the box -> a true lower bound (weak duality), loose
    #     when tiny reduced costs meet the +-delta = 1000 box of the unbounded variables
    zl, zu = np.maximum(z, 0.0), np.minimum(z, 0.0)
    with np.errstate(invalid="ignore"):
        rows = np.sum(np.where(lam > 0, rl * lam, 0.0)) + np.sum(np.where(lam < 0, ru * lam, 0.0))
        strict = rows + np.sum(np.where(zl > 0, lb * zl, 0.0)) + np.sum(np.where(zu < 0, ub * zu, 0.0))
    # (ii) the PDLP convention the engine terminates on: reduced costs are multipliers only where p sits on the
    #     bound, the rest is dual residual (relative to 1 + |c|_2)
    tol = 1e-9 * (1.0 + np.abs(p))
    at_l, at_u = p <= lb + tol, p >= ub - tol
    zl2, zu2 = np.where(at_l, np.maximum(z, 0.0), 0.0), np.where(at_u, np.minimum(z, 0.0), 0.0)
    dres = np.linalg.norm(z - zl2 - zu2) / (1.0 + np.linalg.norm(d["df"]))
    dobj = rows + np.sum(lb * zl2) + np.sum(ub * zu2)
    return pviol / scale, sign_bad, pobj, dobj, dres, strict


@pytest.mark.parametrize("name,eps,engine", [("case1354pegase", 2e-7, 0), ("case13659pegase", 5e-7, 0),
                                             ("case1354pegase", 2e-7, 5)])
def test_full_size_certificate(gpu, name, eps, engine):
    """engine 0: the barrier engine (default); engine 5: the PDHG engines (stream / group hybrid)."""
    from activesetmethods_b200.sublp import SubLp
    pr = problem(name)
    d = _first_linearisation(pr)
    delta = 1000.0
    lp = SubLp(pr.n, pr.m, pr.j_str, pr.x_L, pr.x_U, pr.g_L, pr.g_U, eps_rel=eps, max_iter=8_000_000, engine=engine)
    p, lam, mu_u, mu_l, slack, status = lp.sub_optimize(d["x"], d["f"], d["df"], d["E"], d["dE"], delta, False)
    info = lp.last_info[0]
    assert status == 0, info
    pviol, sign_bad, pobj, dobj, dres, strict = _certificate(pr, d, delta, lp, p, lam)
    gap = abs(pobj - dobj) / (1.0 + abs(pobj) + abs(dobj))
    assert lp.engine_info()["engine"] == (4 if engine == 0 else 2)
    print(f"{name}: {info['iterations']} iterations ({lp.engine_info()}), objective {pobj + d['f']:.6f}, "
          f"relative primal violation {pviol:.2e}, dual residual {dres:.2e}, gap {gap:.2e}; strict lower bound "
          f"{strict + d['f']:.4f} (certified relative gap {abs(pobj - strict) / max(1.0, abs(pobj)):.1e})")
    assert pviol <= 1e-6 and sign_bad <= 1e-9
    assert dres <= 1e-6 and gap <= 1e-6, (pobj, dobj, dres)
    assert strict <= pobj + 1e-6 * abs(pobj) and abs(pobj - strict) <= 1e-2 * max(1.0, abs(pobj))   # weak duality
    assert abs((pobj + d["f"]) - info["objective"]) <= 1e-9 * max(1.0, abs(info["objective"]))
    # multipliers only where the *original* bound is active (subproblem.jl:522-529), with the MOI signs
    assert np.all(mu_u[p < pr.x_U - d["x"]] == 0.0) and np.all(mu_l[p > pr.x_L - d["x"]] == 0.0)
    assert np.all(mu_u <= 0.0) and np.all(mu_l >= 0.0)
    if name == "case1354pegase":
        pat = so.JacobianPattern(pr.m, pr.n, pr.j_str)
        ref = so.SubLp(pat, pr.g_L, pr.g_U, pr.x_L, pr.x_U)
        out = ref.solve(pat.assemble(d["dE"]), d["df"], d["f"], d["E"], d["x"], delta, False)
        assert out[5] == 0
        rel = abs(info["objective"] - ref.last_objective) / max(1.0, abs(ref.last_objective))
        print(f"{name}: objective vs simplex rel {rel:.2e}")
        assert rel <= 1e-6
    lp.close()


@pytest.mark.parametrize("engine", [0, 5])
def test_full_size_batch_matches_single(gpu, engine):
    """A 64-scenario case1354 batch: every scenario's objective equals what a single-LP handle gives for it, and all are
    optimal.  engine 0: barrier engine, scenarios across the lanes of a warp vs. one LP with the terms across the
    threads; engine 5: PDHG streaming kernels, then the group kernel for the stragglers, vs. the group kernel alone."""
    import bench
    from activesetmethods_b200.sublp import SubLp
    net = bench.network("case1354pegase")
    S = 64
    mdl, d = bench.linearise(net, list(range(1, S + 1)))
    lpb = SubLp(mdl.n, mdl.m, mdl.j_str, d["xL"], d["xU"], d["gL"], d["gU"], batch=S, eps_rel=5e-7, max_iter=8_000_000, engine=engine)
    out = lpb.sub_optimize(d["x"], d["f"], d["df"], d["E"], d["dE"], 1000.0, False)
    assert np.all(out[5] == 0)
    objs = np.array([i["objective"] for i in lpb.last_info])
    assert lpb.engine_info()["engine"] == (4 if engine == 0 else 3)          # 3: streamed, then handed over
    for s in (0, 17, 63):
        lp1 = SubLp(mdl.n, mdl.m, mdl.j_str, d["xL"][s], d["xU"][s], d["gL"][s], d["gU"][s], eps_rel=5e-7, max_iter=8_000_000,
                    engine=engine)
        lp1.sub_optimize(d["x"][s], d["f"][s], d["df"][s], d["E"][s], d["dE"][s], 1000.0, False)
        o1 = lp1.last_info[0]["objective"]
        assert lp1.last_info[0]["status"] == 0
        assert abs(o1 - objs[s]) <= 2e-6 * max(1.0, abs(o1)), (s, o1, objs[s])
        lp1.close()
    lpb.close()
