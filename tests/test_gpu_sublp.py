"""GPU parity: the CUDA sub-LP path (through the C ABI) against the CPU oracle on the same linearisations.

Bars (BASELINE.json north_star): CSR assembly bit-exact; per sub-LP objective <= 1e-6 relative, primal and
dual feasibility <= 1e-6; LP status identical."""
import os

import numpy as np
import pytest

from helpers import problem, record_sublps, lp_feasibility, dual_feasibility
from oracle import slp_oracle as so
from test_oracle_pins import case3_network, GOLDEN

pytestmark = pytest.mark.gpu

OBJ_RTOL = 1e-6
FEAS_TOL = 1e-6


def _check_lp(lp, ref, pat, d, n, m):
    """Solve one recorded linearisation on the GPU and compare with the oracle's simplex solve."""
    vals = pat.assemble(d["dE"])
    p, lam, mu_u, mu_l, slack, status = lp.sub_optimize(d["x"], d["f"], d["df"], d["E"], d["dE"], d["delta"], d["fr"])
    rp, ci, v = lp.jacobian_csr()
    assert np.array_equal(v, vals), "device CSR values are not bit-identical to compute_jacobian_matrix"
    ref_out = ref.solve(vals, d["df"], d["f"], d["E"], d["x"], d["delta"], d["fr"])
    assert status == ref_out[5], f"LP status {status} != oracle {ref_out[5]} (fr={d['fr']})"
    info = lp.last_info[0]
    if status != so.OPTIMAL:
        assert not p.any() and not lam.any()
        return info
    K, cost, off, lb, ub, rl, ru = ref.last_lp
    obj_ref = ref.last_objective
    assert abs(info["objective"] - obj_ref) <= OBJ_RTOL * max(1.0, abs(obj_ref)), (info, obj_ref)
    # primal feasibility of the full LP point (p + slacks in the reference's column layout)
    xfull = np.zeros(K.shape[1])
    xfull[:n] = p
    if d["fr"]:
        xfull[ref.s1] = slack[:, 0]
        xfull[ref.s2[ref.two]] = slack[ref.two, 1]
    assert lp_feasibility(K, xfull, lb, ub, rl, ru) <= FEAS_TOL
    assert abs(cost @ xfull + off - info["objective"]) <= 1e-9 * max(1.0, abs(obj_ref))
    # dual feasibility: lambda (range rows already summed) against the J block
    J = pat.matrix(vals)
    c = np.zeros(n) if d["fr"] else d["df"]
    assert dual_feasibility(J, c, p, lam, lb[:n], ub[:n], tol=1e-7) <= FEAS_TOL * max(1.0, np.max(np.abs(c)))
    # multipliers only where the *original* bound is active (subproblem.jl:522-529)
    assert np.all(mu_u[p < (ref.v_ub - d["x"])] == 0.0) and np.all(mu_l[p > (ref.v_lb - d["x"])] == 0.0)
    assert np.all(mu_u <= 0.0) and np.all(mu_l >= 0.0)
    return info


@pytest.mark.parametrize("name", ["toy", "case3", "case9", "case9tr"])
def test_golden_sublps(gpu, name):
    """Committed fixtures (tests/golden/sublp_*.npz): status and objective of every recorded sub-LP."""
    from activesetmethods_b200.sublp import SubLp
    g = np.load(os.path.join(GOLDEN, f"sublp_{name}.npz"))
    from activesetmethods_b200.examples import acopf, small_nlps
    pr = {"toy": small_nlps.ToyNlp, "case3": lambda: acopf.AcopfModel(case3_network()),
          "case9": lambda: acopf.AcopfModel(acopf.case9()), "case9tr": lambda: acopf.AcopfModel(acopf.case9())}[name]()
    lp = SubLp(pr.n, pr.m, pr.j_str, pr.x_L, pr.x_U, pr.g_L, pr.g_U)
    for k in range(len(g["f"])):
        out = lp.sub_optimize(g["x"][k], g["f"][k], g["df"][k], g["E"][k], g["dE"][k], g["delta"][k], bool(g["fr"][k]))
        assert out[5] == int(g["status"][k]), (k, out[5], int(g["status"][k]))
        if out[5] == 0:
            obj = lp.last_info[0]["objective"]
            assert abs(obj - g["objective"][k]) <= OBJ_RTOL * max(1.0, abs(g["objective"][k])), (k, obj)
    lp.close()


@pytest.mark.parametrize("name,alg,limit", [("toy", "Line Search", 14), ("hs071", "Line Search", 6),
                                            ("case9", "Line Search", 6), ("case9", "Trust Region", 6),
                                            ("case118", "Line Search", 3), ("case118", "Trust Region", 3)])
def test_sublp_parity_live(gpu, name, alg, limit):
    from activesetmethods_b200.sublp import SubLp
    pr = problem(name)
    slp, lps = record_sublps(pr, alg, max_iter=40, limit=limit)
    pat = so.JacobianPattern(pr.m, pr.n, pr.j_str)
    lp = SubLp(pr.n, pr.m, pr.j_str, pr.x_L, pr.x_U, pr.g_L, pr.g_U)
    for d in lps:
        ref = so.SubLp(pat, pr.g_L, pr.g_U, pr.x_L, pr.x_U)
        _check_lp(lp, ref, pat, d, pr.n, pr.m)
    lp.close()


def test_numerically_empty_rows_do_not_stall(gpu):
    """Regression (tests/golden/sublp_case118_tinyrow.npz): an SLP iterate whose thermal-limit rows have gradients
    of ~3e-11.  Equilibrating those rows used to scale their right-hand sides by 2e10 and the solve ran into the
    iteration limit; they are now treated as empty rows.  Status and objective against the oracle's simplex."""
    from activesetmethods_b200.sublp import SubLp
    g = np.load(os.path.join(GOLDEN, "sublp_case118_tinyrow.npz"))
    pr = problem("case118")
    pat = so.JacobianPattern(pr.m, pr.n, pr.j_str)
    rowmax = np.asarray(abs(pat.matrix(pat.assemble(g["dE"]))).max(axis=1).todense()).ravel()
    assert np.any((rowmax > 0) & (rowmax < 1e-9))            # the fixture really has such rows
    lp = SubLp(pr.n, pr.m, pr.j_str, pr.x_L, pr.x_U, pr.g_L, pr.g_U, eps_rel=1e-7, max_iter=400000)
    out = lp.sub_optimize(g["x"], float(g["f"]), g["df"], g["E"], g["dE"], float(g["delta"]), bool(g["fr"]))
    info = lp.last_info[0]
    assert out[5] == int(g["status"]) == 0, info
    assert abs(info["objective"] - float(g["objective"])) <= OBJ_RTOL * abs(float(g["objective"]))
    assert info["iterations"] < 200000, info
    lp.close()


@pytest.mark.parametrize("B", [1, 3])
def test_metric_reductions_match_oracle(gpu, B):
    """The device reductions of the iteration — norm_violations (common.jl:75-98, p = 1, 2, inf), KT_residuals
    (:35-44), norm_complementarity (:51-68), the row norms of compute_nu! (slp.jl:54-66), compute_phi (slp.jl:79-115)
    in both phases and compute_derivative (slp.jl:122-147) — against the oracle's functions at random points."""
    from activesetmethods_b200.sublp import SubLp
    pr = problem("case9")
    rng = np.random.default_rng(B)
    x = rng.uniform(np.where(np.isfinite(pr.x_L), pr.x_L, -0.3) - 0.05, np.where(np.isfinite(pr.x_U), pr.x_U, 0.3) + 0.05,
                    (B, pr.n))
    f = np.array([pr.eval_f(xx) for xx in x])
    df = np.array([pr.eval_grad_f(xx, np.zeros(pr.n)) for xx in x])
    E = np.array([pr.eval_g(xx, np.zeros(pr.m)) for xx in x])
    dE = np.array([pr.eval_jac_g(xx, "eval", None, None, np.zeros(len(pr.j_str))) for xx in x])
    lam = rng.standard_normal((B, pr.m)); mu_u = -rng.random((B, pr.n)); mu_l = rng.random((B, pr.n))
    nu = rng.random((B, pr.m)) * 3.0
    pat = so.JacobianPattern(pr.m, pr.n, pr.j_str)
    lp = SubLp(pr.n, pr.m, pr.j_str, pr.x_L, pr.x_U, pr.g_L, pr.g_U, batch=B, eps_rel=1e-8)
    lp._squeeze = False
    close = lambda a, b: np.allclose(np.ravel(a), np.ravel(b), rtol=1e-12, atol=1e-13)   # noqa: E731
    for fr in (False, True):
        lp.update(x, f, df, E, dE, 1000.0, fr)
        for pn in (1, 2, np.inf):
            assert close(lp.norm_violations(None, None, pn),
                         [so.norm_violations(E[s], pr.g_L, pr.g_U, x[s], pr.x_L, pr.x_U, pn) for s in range(B)])
        J = [pat.matrix(pat.assemble(dE[s])) for s in range(B)]
        assert close(lp.kt_residuals(lam, mu_u, mu_l), [so.kt_residuals(df[s], lam[s], mu_u[s], mu_l[s], J[s]) for s in range(B)])
        assert close(lp.norm_complementarity(lam), [so.norm_complementarity(E[s], pr.g_L, pr.g_U, lam[s]) for s in range(B)])
        assert close(lp.row_norms(), [so.row_norms(J[s]) for s in range(B)])
        p, lam_lp, _, _, slack, status = lp.solve_extract()
        if not np.all(np.atleast_1d(status) == 0):
            assert not fr      # a random point may make the normal-phase LP infeasible; the elastic one never is
            continue
        slack = np.asarray(slack).reshape(B, pr.m, 2)
        # the oracle's drivers hold the same formulas: drive one of its objects with this data
        for s in range(B):
            o = so.SlpLS(problem("case9"), so.Parameters())
            o.x, o.f, o.df, o.E, o.dE = x[s].copy(), f[s], df[s].copy(), E[s].copy(), dE[s].copy()
            o.p, o.p_slack, o.nu, o.feasibility_restoration = np.atleast_2d(p)[s], slack[s], nu[s], fr
            o.prim_infeas = 0.37
            alpha = 0.6
            Et = pr.eval_g(x[s] + alpha * o.p, np.zeros(pr.m))
            base = 0.37 if fr else pr.eval_f(x[s] + alpha * o.p)
            Ets = np.tile(Et, (B, 1)); bases = np.full(B, base); alphas = np.full(B, alpha)
            got = np.atleast_1d(lp.merit_phi(bases, Ets, nu, alphas, fr))[s]
            assert abs(got - o.compute_phi(x[s], alpha, o.p)) <= 1e-10 * max(1.0, abs(got)), (fr, s)
            got0 = np.atleast_1d(lp.merit_phi(np.full(B, 0.37 if fr else f[s]), None, nu, np.zeros(B), fr))[s]
            assert abs(got0 - o.compute_phi(x[s], 0.0, o.p)) <= 1e-10 * max(1.0, abs(got0)), (fr, s)
            gotd = np.atleast_1d(lp.merit_derivative(nu, fr))[s]
            assert abs(gotd - o.compute_derivative()) <= 1e-9 * max(1.0, abs(gotd)), (fr, s)
    lp.close()


def test_assembly_bit_exact_with_duplicates(gpu):
    """common.jl:12-20 with duplicate COO entries, wide dynamic range, signed zeros; batch of 40 scenarios."""
    from activesetmethods_b200.examples import small_nlps
    from activesetmethods_b200.sublp import SubLp
    pr = small_nlps.RandomNlp()
    pat = so.JacobianPattern(pr.m, pr.n, pr.j_str)
    assert pat.nnz < len(pr.j_str)
    rng = np.random.default_rng(0)
    for B in (1, 40):
        lp = SubLp(pr.n, pr.m, pr.j_str, pr.x_L, pr.x_U, pr.g_L, pr.g_U, batch=B)
        nz = len(pr.j_str)
        dE = rng.standard_normal((B, nz)) * 10.0 ** rng.integers(-12, 12, (B, nz))
        dE[:, ::7] = -0.0
        x = np.tile(pr.x0, (B, 1))
        lp.update(x, np.zeros(B), np.zeros((B, pr.n)), np.zeros((B, pr.m)), dE, 1.0, False)
        for s in range(B):
            rp, ci, v = lp.jacobian_csr(s)
            ref = pat.assemble(dE[s])
            assert np.array_equal(rp, pat.row_ptr) and np.array_equal(ci, pat.cols)
            assert np.array_equal(v, ref) and np.array_equal(np.signbit(v), np.signbit(ref))
        lp.close()


def test_empty_and_error_paths(gpu):
    from activesetmethods_b200 import capi
    from activesetmethods_b200.sublp import SubLp
    # free row -> rejected like the reference builder (subproblem.jl:143-197 pushes no row)
    with pytest.raises(capi.AsmError) as e:
        SubLp(2, 1, np.array([[1, 1]]), [-1, -1], [1, 1], [-np.inf], [np.inf])
    assert e.value.code == capi.E_FREE_ROW
    # out-of-range pattern entry
    with pytest.raises(capi.AsmError) as e:
        SubLp(2, 1, np.array([[2, 1]]), [-1, -1], [1, 1], [0.0], [1.0])
    assert e.value.code == capi.E_INVALID
    # solve before update
    lp = SubLp(2, 1, np.array([[1, 1], [1, 2]]), [-1, -1], [1, 1], [0.0], [1.0])
    with pytest.raises(capi.AsmError) as e:
        lp.solve()
    assert e.value.code == capi.E_STATE
    # m = 0: a pure box LP  min df'p, |p| <= delta
    lp0 = SubLp(3, 0, np.zeros((0, 2), dtype=np.int64), [-5, -5, -5], [5, 5, 5], [], [])
    p, lam, mu_u, mu_l, slack, status = lp0.sub_optimize(np.zeros(3), 0.0, np.array([1.0, -2.0, 0.0]), [], [], 0.5)
    assert status == 0 and np.allclose(p[:2], [-0.5, 0.5]) and abs(lp0.last_info[0]["objective"] + 1.5) < 1e-7


def test_batch_matches_single(gpu):
    """A batch of perturbed case9 linearisations gives, per scenario, what the single-LP path gives."""
    from activesetmethods_b200.examples import acopf
    from activesetmethods_b200.sublp import SubLp
    net = acopf.case9()
    B = 5
    mdls = [acopf.AcopfModel(acopf.perturb_loads(net, s + 1)) for s in range(B)]
    m0 = mdls[0]
    gL = np.array([m.g_L for m in mdls]); gU = np.array([m.g_U for m in mdls])
    xL = np.array([m.x_L for m in mdls]); xU = np.array([m.x_U for m in mdls])
    x = np.array([np.clip(m.x0, m.x_L, m.x_U) for m in mdls])
    f = np.array([m.eval_f(xx) for m, xx in zip(mdls, x)])
    df = np.array([m.eval_grad_f(xx, np.zeros(m.n)) for m, xx in zip(mdls, x)])
    E = np.array([m.eval_g(xx, np.zeros(m.m)) for m, xx in zip(mdls, x)])
    dE = np.array([m.eval_jac_g(xx, "eval", None, None, np.zeros(m.nnz)) for m, xx in zip(mdls, x)])
    lpb = SubLp(m0.n, m0.m, m0.j_str, xL, xU, gL, gU, batch=B)
    pb, lamb, _, _, _, stb = lpb.sub_optimize(x, f, df, E, dE, 1000.0, False)
    pat = so.JacobianPattern(m0.m, m0.n, m0.j_str)
    for s in range(B):
        ref = so.SubLp(pat, gL[s], gU[s], xL[s], xU[s])
        ref.solve(pat.assemble(dE[s]), df[s], f[s], E[s], x[s], 1000.0, False)
        assert stb[s] == 0
        obj = lpb.last_info[s]["objective"]
        assert abs(obj - ref.last_objective) <= OBJ_RTOL * max(1.0, abs(ref.last_objective)), (s, obj)
    lpb.close()


def test_large_batch_with_compaction(gpu):
    """160 perturbed case9 scenarios: wide enough (>= 128) for the streaming engine to compact the running LPs into
    a narrower working set as they converge and to hand the stragglers to the group kernel.  Every scenario must
    come back in its own slot: objective against the oracle for scenarios spread over the batch, all optimal, primal
    steps feasible for their own (per-scenario) bounds."""
    from activesetmethods_b200.examples import acopf
    from activesetmethods_b200.sublp import SubLp
    net = acopf.case9()
    B = 160
    mdls = [acopf.AcopfModel(acopf.perturb_loads(net, s + 1)) for s in range(B)]
    m0 = mdls[0]
    gL = np.array([m.g_L for m in mdls]); gU = np.array([m.g_U for m in mdls])
    xL = np.array([m.x_L for m in mdls]); xU = np.array([m.x_U for m in mdls])
    x = np.array([np.clip(m.x0, m.x_L, m.x_U) for m in mdls])
    f = np.array([m.eval_f(xx) for m, xx in zip(mdls, x)])
    df = np.array([m.eval_grad_f(xx, np.zeros(m.n)) for m, xx in zip(mdls, x)])
    E = np.array([m.eval_g(xx, np.zeros(m.m)) for m, xx in zip(mdls, x)])
    dE = np.array([m.eval_jac_g(xx, "eval", None, None, np.zeros(m.nnz)) for m, xx in zip(mdls, x)])
    lpb = SubLp(m0.n, m0.m, m0.j_str, xL, xU, gL, gU, batch=B, eps_rel=1e-7, engine=5)     # PDHG hybrid
    pb, lamb, _, _, _, stb = lpb.sub_optimize(x, f, df, E, dE, 1000.0, False)
    assert np.all(stb == 0)
    objs = np.array([i["objective"] for i in lpb.last_info])
    its = np.array([i["iterations"] for i in lpb.last_info])
    assert its.max() > 2 * np.median(its) or its.min() < its.max()      # they do not all stop together
    pat = so.JacobianPattern(m0.m, m0.n, m0.j_str)
    for s in (0, 1, 63, 64, 77, 128, 159):
        ref = so.SubLp(pat, gL[s], gU[s], xL[s], xU[s])
        ref.solve(pat.assemble(dE[s]), df[s], f[s], E[s], x[s], 1000.0, False)
        assert abs(objs[s] - ref.last_objective) <= OBJ_RTOL * max(1.0, abs(ref.last_objective)), (s, objs[s])
        K, cost, off, lb, ub, rl, ru = ref.last_lp
        xfull = np.zeros(K.shape[1]); xfull[:m0.n] = pb[s]
        assert lp_feasibility(K, xfull, lb, ub, rl, ru) <= FEAS_TOL, s
    lpb.close()


def test_batch_in_feasibility_restoration(gpu):
    """A batch of restoration-phase LPs (subproblem.jl:250-382: elastic rows, slack columns, shifted right-hand
    sides) of the toy NLP at scattered points: objective and slack sums per scenario against the oracle."""
    from activesetmethods_b200.examples import small_nlps
    from activesetmethods_b200.sublp import SubLp
    pr = small_nlps.ToyNlp()
    rng = np.random.default_rng(11)
    B = 40
    x = rng.uniform(-1.5, 1.5, (B, pr.n))
    f = np.array([pr.eval_f(xx) for xx in x])
    df = np.array([pr.eval_grad_f(xx, np.zeros(pr.n)) for xx in x])
    E = np.array([pr.eval_g(xx, np.zeros(pr.m)) for xx in x])
    dE = np.array([pr.eval_jac_g(xx, "eval", None, None, np.zeros(len(pr.j_str))) for xx in x])
    lpb = SubLp(pr.n, pr.m, pr.j_str, pr.x_L, pr.x_U, pr.g_L, pr.g_U, batch=B, eps_rel=1e-8)
    p, lam, mu_u, mu_l, slack, status = lpb.sub_optimize(x, f, df, E, dE, 1000.0, True)
    pat = so.JacobianPattern(pr.m, pr.n, pr.j_str)
    for s in range(B):
        ref = so.SubLp(pat, pr.g_L, pr.g_U, pr.x_L, pr.x_U)
        out = ref.solve(pat.assemble(dE[s]), df[s], f[s], E[s], x[s], 1000.0, True)
        assert status[s] == out[5] == 0, (s, status[s], out[5])
        obj = lpb.last_info[s]["objective"]
        assert abs(obj - ref.last_objective) <= OBJ_RTOL * max(1.0, abs(ref.last_objective)), (s, obj, ref.last_objective)
        assert abs(slack[s].sum() - obj) <= 1e-6 * max(1.0, abs(obj))      # restoration objective = sum of slacks
    lpb.close()


def test_generic_lp_against_highs(gpu):
    """B200LP as a general external LP optimizer: random feasible bounded LPs against HiGHS."""
    import scipy.sparse as sp
    from activesetmethods_b200.sublp import B200LP
    rng = np.random.default_rng(5)
    for trial in range(3):
        n, m = 30 + 10 * trial, 20 + 5 * trial
        K = sp.random(m, n, density=0.2, random_state=trial, format="csr")
        K.data = rng.standard_normal(len(K.data))
        x0 = rng.uniform(-1, 1, n)
        Kx = K @ x0
        rl = Kx - rng.uniform(0, 1, m); ru = Kx + rng.uniform(0, 1, m)
        rl[::3] = ru[::3] = Kx[::3]
        rl[1::5] = -np.inf
        lb = np.full(n, -2.0); ub = np.full(n, 2.0)
        c = rng.standard_normal(n)
        eng = so.HighsLp()
        st, xr, yr, dr, obj = eng.solve(K.tocsc(), c, 0.25, lb, ub, rl, ru)
        assert st == so.OPTIMAL
        lp = B200LP(n, m, K.indptr, K.indices)
        lp.set_matrix_values(K.data); lp.set_objective(c, 0.25); lp.set_col_bounds(lb, ub); lp.set_row_bounds(rl, ru)
        info = lp.optimize()[0]
        assert info["status"] == 0
        assert abs(info["objective"] - obj) <= OBJ_RTOL * max(1.0, abs(obj))
        assert lp_feasibility(K, lp.primal(), lb, ub, rl, ru) <= FEAS_TOL
        lp.close()


def test_set_active_masks_scenarios(gpu):
    """asm_slp_set_active: masked scenarios cost nothing and come back SKIPPED with zeroed outputs, the others are
    solved exactly as in the unmasked batch (what the lock-step SLP driver does for finished / other-phase scenarios)."""
    from activesetmethods_b200.examples import acopf
    from activesetmethods_b200.sublp import SubLp
    net = acopf.case9()
    B = 40
    mdls = [acopf.AcopfModel(acopf.perturb_loads(net, s + 1)) for s in range(B)]
    m0 = mdls[0]
    arr = lambda f: np.array([f(m) for m in mdls])   # noqa: E731
    x = arr(lambda m: np.clip(m.x0, m.x_L, m.x_U))
    args = (x, np.array([m.eval_f(xx) for m, xx in zip(mdls, x)]),
            np.array([m.eval_grad_f(xx, np.zeros(m.n)) for m, xx in zip(mdls, x)]),
            np.array([m.eval_g(xx, np.zeros(m.m)) for m, xx in zip(mdls, x)]),
            np.array([m.eval_jac_g(xx, "eval", None, None, np.zeros(m.nnz)) for m, xx in zip(mdls, x)]), 1000.0, False)
    lp = SubLp(m0.n, m0.m, m0.j_str, arr(lambda m: m.x_L), arr(lambda m: m.x_U), arr(lambda m: m.g_L),
               arr(lambda m: m.g_U), batch=B)
    full = lp.sub_optimize(*args)
    obj_full = np.array([i["objective"] for i in lp.last_info])
    mask = np.zeros(B, dtype=bool)
    mask[[0, 3, 17, 39]] = True
    lp.set_active(mask)
    part = lp.sub_optimize(*args)
    st = np.array([i["status"] for i in lp.last_info])
    assert np.all(st[mask] == 0) and np.all(st[~mask] == 5)
    assert np.all(part[5][~mask] == 5) and not part[0][~mask].any() and not part[1][~mask].any()
    obj_part = np.array([i["objective"] for i in lp.last_info])
    assert np.allclose(obj_part[mask], obj_full[mask], rtol=1e-12, atol=0)
    assert np.allclose(part[0][mask], full[0][mask], rtol=0, atol=1e-9)
    lp.set_active(None)
    again = lp.sub_optimize(*args)
    assert np.all(again[5] == 0)
    lp.close()


def test_pageable_and_pinned_inputs_agree(gpu):
    """Host arrays go straight to the device when the caller pinned them and through the handle's pinned ring when they
    are pageable (util.cuh PinnedRing): same bits either way, also for inputs larger than one ring chunk."""
    import torch
    from activesetmethods_b200.sublp import SubLp
    pr = problem("case118")
    B = 96                                            # dE: 96 x 6297 x 8 B = 4.8 MB; two arrays exceed a ring chunk together
    x = np.clip(pr.x0, pr.x_L, pr.x_U)
    one = dict(x=x, f=pr.eval_f(x), df=pr.eval_grad_f(x, np.zeros(pr.n)), E=pr.eval_g(x, np.zeros(pr.m)),
               dE=pr.eval_jac_g(x, "eval", None, None, np.zeros(len(pr.j_str))))
    rng = np.random.default_rng(0)
    scale = 1.0 + 1e-3 * rng.standard_normal((B, 1))
    big = {k: np.ascontiguousarray(np.repeat(np.atleast_2d(v), B, axis=0) * (scale if k in ("df", "dE") else 1.0))
           for k, v in one.items() if k != "f"}
    f = np.full(B, one["f"])
    lp = SubLp(pr.n, pr.m, pr.j_str, pr.x_L, pr.x_U, pr.g_L, pr.g_U, batch=B)
    lp.update(big["x"], f, big["df"], big["E"], big["dE"], 1000.0, False)
    v_pageable = lp.jacobian_csr(7)[2].copy()
    pinned = {k: torch.from_numpy(v.copy()).pin_memory().numpy() for k, v in big.items()}
    lp.update(pinned["x"], f, pinned["df"], pinned["E"], pinned["dE"], 1000.0, False)
    v_pinned = lp.jacobian_csr(7)[2].copy()
    assert np.array_equal(v_pageable, v_pinned)
    pat = so.JacobianPattern(pr.m, pr.n, pr.j_str)
    assert np.array_equal(v_pinned, pat.assemble(big["dE"][7]))
    lp.close()
