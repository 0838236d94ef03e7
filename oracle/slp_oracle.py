"""CPU ORACLE — test infrastructure only.  Nothing in ``activesetmethods_b200/`` may import this file.

A restatement, in numpy / plain Python, of the reference's per-iteration hot path
(exanauts/ActiveSetMethods v0.1.0):

  ====================================  ===============================================================
  here                                  reference (relative to /root/reference)
  ====================================  ===============================================================
  ``JacobianPattern`` / ``.assemble``   ``src/algorithms/common.jl:12-20``  (ordered duplicate sum)
  ``SubLp.__init__``                    ``src/algorithms/subproblem.jl:51-215``  (create_model!)
  ``SubLp.solve``                       ``src/algorithms/subproblem.jl:229-542`` (sub_optimize!)
  ``kt_residuals``                      ``src/algorithms/common.jl:35-44``
  ``norm_complementarity``              ``src/algorithms/common.jl:51-68``
  ``norm_violations``                   ``src/algorithms/common.jl:75-98``
  ``Slp.compute_phi``                   ``src/algorithms/slp.jl:79-115``
  ``Slp.compute_derivative``            ``src/algorithms/slp.jl:122-147``
  ``SlpLS``                             ``src/algorithms/slp_line_search.jl:4-261``
  ``SlpTR``                             ``src/algorithms/slp_trust_region.jl:10-251`` (+ ``slp.jl:54-66``)
  ``Parameters``                        ``src/parameters.jl:1-29``
  ====================================  ===============================================================

LP engine.  The reference hands the LP to GLPK (GLPK.jl 0.13.0 / GLPK_jll 4.64.0, pinned in
``examples/Manifest.toml:98-108``) through MOI.  GLPK is not in this image; the stand-in vertex solver is
HiGHS (simplex) as bundled with SciPy 1.18 (``scipy.optimize._highspy._core``), driven with a persistent
model and basis warm start — the analogue of the persistent ``glp_prob`` the reference keeps
(``src/algorithms/slp.jl:24,39``).

PARITY PINS.  No reference test pins a result at the sub-LP boundary (``test/unittests.jl`` is empty), so at
that boundary parity is *unpinned*; what *is* pinned, and what ``tests/test_oracle_pins.py`` checks this oracle
against, are the reference's end-to-end known answers: the toy NLP → X = Y = -1, LOCALLY_SOLVED
(``test/runtests.jl:9-14``), ACP-OPF ``case3.m`` objective 5906.87949 (``test/runtests.jl:18-21``), plus the
derivable facts "first toy LP is INFEASIBLE, its FR optimum is 4.0" and the public optima MATPOWER case9
5296.69 and hs071 17.0140173.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

INF = math.inf

# LP statuses (the subset of MOI.TerminationStatus the reference branches on, subproblem.jl:500-539)
OPTIMAL, INFEASIBLE, DUAL_INFEASIBLE, OTHER_ERROR = 0, 1, 2, 3


@dataclass
class Parameters:
    """src/parameters.jl:1-29 (fields read on the path)."""
    algorithm: str = "Line Search"
    OutputFlag: int = 0
    tol_direction: float = 1.0e-6
    tol_residual: float = 0.01
    tol_infeas: float = 0.01
    max_iter: int = 1000
    eta: float = 0.4
    tau: float = 0.9
    min_alpha: float = 1.0e-6
    tr_size: float = 0.4


# ----------------------------------------------------------------------------------------------------------
# common.jl
# ----------------------------------------------------------------------------------------------------------
class JacobianPattern:
    """Fixed sparsity of ``J`` from the COO list ``j_str`` (1-based pairs, duplicates allowed).

    ``assemble(dE)`` reproduces ``compute_jacobian_matrix`` (common.jl:12-20): every stored entry is
    ``((0.0 + v1) + v2) + ...`` with the duplicates taken in ``j_str`` order.  The pattern keeps explicit
    zeros."""

    def __init__(self, m: int, n: int, j_str: np.ndarray):
        j_str = np.asarray(j_str, dtype=np.int64).reshape(-1, 2)
        self.m, self.n = m, n
        r = j_str[:, 0] - 1
        c = j_str[:, 1] - 1
        if len(r) and (r.min() < 0 or r.max() >= m or c.min() < 0 or c.max() >= n):
            raise IndexError("j_str entry out of range")
        key = r * n + c
        order = np.argsort(key, kind="stable")
        ks = key[order]
        first = np.ones(len(ks), dtype=bool)
        first[1:] = ks[1:] != ks[:-1]
        slot_sorted = np.cumsum(first) - 1
        self.nnz = int(slot_sorted[-1] + 1) if len(ks) else 0
        self.slot = np.empty(len(ks), dtype=np.int64)         # COO entry -> CSR slot
        self.slot[order] = slot_sorted
        start = np.nonzero(first)[0]
        rank_sorted = np.arange(len(ks)) - start[slot_sorted] if len(ks) else np.zeros(0, dtype=np.int64)
        self.rank = np.empty(len(ks), dtype=np.int64)         # occurrence number within the slot
        self.rank[order] = rank_sorted
        self.max_rank = int(rank_sorted.max()) if len(ks) else -1
        uk = ks[first]
        self.rows = (uk // n).astype(np.int64)
        self.cols = (uk % n).astype(np.int64)
        self.row_ptr = np.zeros(m + 1, dtype=np.int64)
        np.add.at(self.row_ptr, self.rows + 1, 1)
        self.row_ptr = np.cumsum(self.row_ptr)

    def assemble(self, dE: np.ndarray) -> np.ndarray:
        vals = np.zeros(self.nnz)
        for k in range(self.max_rank + 1):
            sel = self.rank == k
            vals[self.slot[sel]] = vals[self.slot[sel]] + dE[sel]    # one touch per slot per round
        return vals

    def matrix(self, vals: np.ndarray) -> sp.csr_matrix:
        return sp.csr_matrix((vals, self.cols, self.row_ptr), shape=(self.m, self.n))


def assemble_reference_loop(m, n, j_str, dE):
    """Literal scalar restatement of common.jl:15-18 (dict-of-keys instead of Julia's CSC insert) — used
    on small inputs to pin ``JacobianPattern.assemble``."""
    J = {}
    for i in range(len(j_str)):
        k = (int(j_str[i][0]), int(j_str[i][1]))
        J[k] = J.get(k, 0.0) + float(dE[i])
    return J


def row_norms(J: sp.csr_matrix) -> np.ndarray:
    return np.sqrt(np.asarray(J.multiply(J).sum(axis=1)).ravel())


def kt_residuals(df, lam, mult_x_U, mult_x_L, J: sp.csr_matrix) -> float:
    """common.jl:35-44."""
    kt = np.linalg.norm(df - J.T @ lam - mult_x_U - mult_x_L)
    scalar = max(1.0, float(np.linalg.norm(df)))
    if J.shape[0]:
        scalar = max(scalar, float(np.max(np.abs(lam) * row_norms(J))))
    return float(kt / scalar)


def norm_complementarity(E, g_L, g_U, lam, p=INF) -> float:
    """common.jl:51-68 (x, bounds and bound multipliers are accepted but unused there)."""
    ineq = g_L != g_U
    compl = np.zeros(len(E))
    with np.errstate(invalid="ignore"):
        compl[ineq] = np.minimum(E[ineq] - g_L[ineq], g_U[ineq] - E[ineq]) * lam[ineq]
    denom = float(np.sum(lam[ineq] ** 2))
    nrm = float(np.max(np.abs(compl))) if (p == INF and len(compl)) else float(np.linalg.norm(compl, p)) if len(compl) else 0.0
    return nrm / (1.0 + math.sqrt(denom))


def norm_violations(E, g_L, g_U, x, x_L, x_U, p=1) -> float:
    """common.jl:75-98."""
    viol = np.concatenate([
        np.where(E > g_U, E - g_U, np.where(E < g_L, g_L - E, 0.0)),
        np.where(x > x_U, x - x_U, np.where(x < x_L, x_L - x, 0.0)),
    ])
    if len(viol) == 0:
        return 0.0
    return float(np.max(np.abs(viol))) if p == INF else float(np.linalg.norm(viol, p))


# ----------------------------------------------------------------------------------------------------------
# LP engine: HiGHS simplex, persistent model, basis warm start
# ----------------------------------------------------------------------------------------------------------
class HighsLp:
    def __init__(self, threads: int = 1):
        import scipy.optimize._highspy._core as hc
        self.hc = hc
        self.h = hc._Highs()
        self.h.setOptionValue("output_flag", False)
        self.h.setOptionValue("solver", "simplex")
        self.h.setOptionValue("threads", threads)
        self.h.setOptionValue("primal_feasibility_tolerance", 1e-9)
        self.h.setOptionValue("dual_feasibility_tolerance", 1e-9)
        self._have_basis = False

    def solve(self, K: sp.csc_matrix, c, c0, lb, ub, rl, ru):
        """min c'x + c0  s.t.  rl <= Kx <= ru, lb <= x <= ub.  Returns (status, x, row_dual, col_dual, obj)."""
        hc, h = self.hc, self.h
        basis = h.getBasis() if self._have_basis else None
        lp = hc.HighsLp()
        m, n = K.shape
        lp.num_col_, lp.num_row_ = n, m
        lp.col_cost_ = np.asarray(c, dtype=float)
        lp.col_lower_ = np.asarray(lb, dtype=float)
        lp.col_upper_ = np.asarray(ub, dtype=float)
        lp.row_lower_ = np.asarray(rl, dtype=float)
        lp.row_upper_ = np.asarray(ru, dtype=float)
        lp.offset_ = float(c0)
        K = K.tocsc()
        lp.a_matrix_.format_ = hc.MatrixFormat.kColwise
        lp.a_matrix_.start_ = K.indptr.astype(np.int32)
        lp.a_matrix_.index_ = K.indices.astype(np.int32)
        lp.a_matrix_.value_ = K.data.astype(float)
        h.passModel(lp)
        if basis is not None and len(basis.col_status) == n and len(basis.row_status) == m:
            h.setBasis(basis)
        h.run()
        ms = h.getModelStatus()
        if ms == hc.HighsModelStatus.kOptimal:
            sol = h.getSolution()
            self._have_basis = True
            return (OPTIMAL, np.array(sol.col_value), np.array(sol.row_dual), np.array(sol.col_dual),
                    float(h.getObjectiveValue()))
        self._have_basis = False
        if ms == hc.HighsModelStatus.kInfeasible:
            return INFEASIBLE, None, None, None, None
        if ms in (hc.HighsModelStatus.kUnbounded, hc.HighsModelStatus.kUnboundedOrInfeasible):
            return DUAL_INFEASIBLE, None, None, None, None
        return OTHER_ERROR, None, None, None, None


# ----------------------------------------------------------------------------------------------------------
# subproblem.jl
# ----------------------------------------------------------------------------------------------------------
class SubLp:
    """``QpModel`` + ``create_model!`` + ``sub_optimize!`` (subproblem.jl:16-542), LP algebra of
    SURVEY.md App. A.  Columns: p[0:n], then per row i one slack, or two when both row bounds are finite.
    Rows: m rows, then one extra ``<=`` row per range row (``adj``)."""

    def __init__(self, pattern: JacobianPattern, c_lb, c_ub, v_lb, v_ub, engine=None):
        self.pat = pattern
        n, m = pattern.n, pattern.m
        self.n, self.m = n, m
        self.c_lb, self.c_ub = np.asarray(c_lb, float), np.asarray(c_ub, float)
        self.v_lb, self.v_ub = np.asarray(v_lb, float), np.asarray(v_ub, float)
        lb_f = self.c_lb > -INF
        ub_f = self.c_ub < INF
        if np.any(~lb_f & ~ub_f):
            # subproblem.jl:143-197 pushes no row for a free row and later indexing breaks
            raise ValueError("free rows are not supported by the reference sub-LP builder")
        self.two = lb_f & ub_f                                            # :87
        self.is_eq = self.c_lb == self.c_ub                               # :143
        self.is_rng = (self.c_lb != -INF) & (self.c_ub != INF) & (self.c_lb < self.c_ub) & ~self.is_eq   # :158
        self.is_lo = ~self.is_eq & ~self.is_rng & (self.c_lb != -INF)     # :173
        self.is_up = ~self.is_eq & ~self.is_rng & ~self.is_lo & (self.c_ub != INF)   # :185
        self.adj = np.nonzero(self.is_rng)[0]
        nsl = 1 + self.two.astype(np.int64)
        self.s1 = n + np.concatenate([[0], np.cumsum(nsl)[:-1]]) if m else np.zeros(0, dtype=np.int64)
        self.s2 = np.where(self.two, self.s1 + 1, -1)
        self.ncol = n + int(nsl.sum())
        self.nrow = m + len(self.adj)
        # slack block of the constraint matrix (fixed)
        r, c, v = [], [], []
        idx = np.arange(m)
        e = self.is_eq
        r += [idx[e], idx[e]]; c += [self.s1[e], self.s2[e]]; v += [np.ones(e.sum()), -np.ones(e.sum())]
        g = self.is_rng | self.is_lo
        r += [idx[g]]; c += [self.s1[g]]; v += [np.ones(g.sum())]
        u = self.is_up
        r += [idx[u]]; c += [self.s1[u]]; v += [-np.ones(u.sum())]
        r += [m + np.arange(len(self.adj))]; c += [self.s2[self.adj]]; v += [-np.ones(len(self.adj))]
        self.S = sp.csr_matrix((np.concatenate(v), (np.concatenate(r), np.concatenate(c))),
                               shape=(self.nrow, self.ncol))
        self.engine = engine if engine is not None else HighsLp()
        self.last_objective = None
        self.last_lp = None

    def build(self, vals, c, c0, b, x_k, delta, feasibility):
        """The data push of subproblem.jl:248-484; returns (K, cost, c0, lb, ub, rl, ru)."""
        n, m = self.n, self.m
        J = self.pat.matrix(vals)
        JJ = sp.vstack([J, J[self.adj]], format="csr") if len(self.adj) else J
        K = (sp.hstack([JJ, sp.csr_matrix((self.nrow, self.ncol - n))], format="csr") + self.S).tocsr()
        cost = np.zeros(self.ncol)
        lb = np.empty(self.ncol)
        ub = np.empty(self.ncol)
        bb = np.array(b, dtype=float)
        if feasibility:                                                   # :250-382
            off = 0.0
            cost[n:] = 1.0
            viol = np.where(b > self.c_ub, self.c_ub - b, np.where(b < self.c_lb, self.c_lb - b, 0.0))   # :289-294
            bb = bb - np.abs(viol)                                        # :295
            two = self.two
            neg = viol < 0
            lb[self.s1[two & neg]] = 0.0
            lb[self.s2[two & neg]] = viol[two & neg]
            lb[self.s1[two & ~neg]] = -viol[two & ~neg]
            lb[self.s2[two & ~neg]] = 0.0
            lb[self.s1[~two]] = -np.abs(viol[~two])
            ub[n:] = INF
        else:                                                             # :383-424
            off = float(c0)
            cost[:n] = c
            lb[n:] = 0.0
            ub[n:] = 0.0
        ub[:n] = np.minimum(delta, self.v_ub - x_k)                       # :427-434
        lb[:n] = np.maximum(-delta, self.v_lb - x_k)
        lb[:n] = np.where(lb[:n] == 0.0, 0.0, lb[:n])                     # tol_error = 0 only normalises -0.0
        ub[:n] = np.where(ub[:n] == 0.0, 0.0, ub[:n])
        c_ub = self.c_ub - bb                                             # :461-484
        c_lb = self.c_lb - bb
        rl = np.full(self.nrow, -INF)
        ru = np.full(self.nrow, INF)
        e = self.is_eq
        rl[:m][e] = c_lb[e]; ru[:m][e] = c_lb[e]
        g = self.is_rng | self.is_lo
        rl[:m][g] = c_lb[g]
        u = self.is_up
        ru[:m][u] = c_ub[u]
        ru[m:] = c_ub[self.adj]
        return K, cost, off, lb, ub, rl, ru

    def solve(self, vals, c, c0, b, x_k, delta, feasibility=False):
        """Returns (Xsol, lambda, mult_x_U, mult_x_L, p_slack, status) as subproblem.jl:541;
        ``p_slack`` is an (m, 2) array (second column 0 where a row has one slack) or None."""
        n, m = self.n, self.m
        K, cost, off, lb, ub, rl, ru = self.build(vals, c, c0, b, x_k, delta, feasibility)
        self.last_lp = (K, cost, off, lb, ub, rl, ru)
        status, x, rdual, cdual, obj = self.engine.solve(K, cost, off, lb, ub, rl, ru)
        self.last_objective = obj
        if status == OPTIMAL:
            Xsol = x[:n].copy()
            p_slack = np.zeros((m, 2))
            p_slack[:, 0] = x[self.s1]
            p_slack[self.two, 1] = x[self.s2[self.two]]
            lam = rdual[:m].copy()                                        # :510-515
            lam[self.adj] += rdual[m:]
            mult_x_U = np.minimum(cdual[:n], 0.0)                         # dual of p <= ub (LessThan: <= 0)
            mult_x_L = np.maximum(cdual[:n], 0.0)                         # dual of p >= lb (GreaterThan: >= 0)
            mult_x_U[Xsol < self.v_ub - x_k] = 0.0                        # :522-529
            mult_x_L[Xsol > self.v_lb - x_k] = 0.0
            return Xsol, lam, mult_x_U, mult_x_L, p_slack, OPTIMAL
        if status == INFEASIBLE:                                          # :532-536
            return np.zeros(n), np.zeros(m), np.zeros(n), np.zeros(n), None, INFEASIBLE
        return np.zeros(n), np.zeros(m), np.zeros(n), np.zeros(n), None, status


# ----------------------------------------------------------------------------------------------------------
# slp.jl + drivers
# ----------------------------------------------------------------------------------------------------------
class _Slp:
    def __init__(self, problem, options: Parameters, engine_factory=None):
        self.problem = problem
        self.options = options
        n, m = problem.n, problem.m
        self.x = np.array(problem.x0, dtype=float)
        self.p = np.zeros(n)
        self.p_slack = None
        self.lam = np.zeros(m)
        self.mult_x_L = np.zeros(n)
        self.mult_x_U = np.zeros(n)
        self.f = 0.0
        self.df = np.zeros(n)
        self.E = np.zeros(m)
        self.dE = np.zeros(len(problem.j_str))
        self.phi = INF
        self.nu = np.zeros(m)
        self.prim_infeas = INF
        self.dual_infeas = INF
        self.compl = INF
        self.feasibility_restoration = False
        self.iter = 1
        self.ret = -5
        self.pattern = JacobianPattern(m, n, problem.j_str)
        self.optimizer = None
        self.engine_factory = engine_factory
        self.lp_log = []           # (status, objective, fr) per LP, for parity tests
        self.record = None         # optional callable(slp, lp_inputs)

    # slp.jl:186-191
    def eval_functions(self):
        pr = self.problem
        self.f = pr.eval_f(self.x)
        pr.eval_grad_f(self.x, self.df)
        pr.eval_g(self.x, self.E)
        pr.eval_jac_g(self.x, "eval", None, None, self.dE)

    def jac(self):
        return self.pattern.matrix(self.pattern.assemble(self.dE))

    # slp.jl:23-47
    def sub_optimize(self, delta=1000.0):
        pr = self.problem
        if self.optimizer is None:
            eng = self.engine_factory() if self.engine_factory else None
            self.optimizer = SubLp(self.pattern, pr.g_L, pr.g_U, pr.x_L, pr.x_U, eng)
        vals = self.pattern.assemble(self.dE)
        if self.record is not None:
            self.record(self, dict(x=self.x.copy(), f=self.f, df=self.df.copy(), E=self.E.copy(),
                                   dE=self.dE.copy(), delta=delta, fr=self.feasibility_restoration))
        out = self.optimizer.solve(vals, self.df, self.f, self.E, self.x, delta, self.feasibility_restoration)
        self.lp_log.append((out[5], self.optimizer.last_objective, self.feasibility_restoration))
        return out

    def clip_start(self):
        pr = self.problem
        for i in range(pr.n):                                             # slp_line_search.jl:98-105
            if pr.x_L[i] > -INF:
                self.x[i] = max(self.x[i], pr.x_L[i])
            if pr.x_U[i] > -INF:                                          # sic (App. C-5)
                self.x[i] = min(self.x[i], pr.x_U[i])

    def _row_viol(self, E):
        pr = self.problem
        return np.maximum(0.0, np.maximum(E - pr.g_U, pr.g_L - E))

    # slp.jl:79-115
    def compute_phi(self, x, alpha, p):
        pr = self.problem
        xp = x + alpha * p
        E = self.E if alpha == 0.0 else pr.eval_g(xp, np.zeros(pr.m))
        if self.feasibility_restoration:
            ps = self.p_slack
            phi = self.prim_infeas
            phi += alpha * float(np.sum(ps))                              # :86-88 (sum order differs: Dict)
            viol = self._row_viol(self.E)                                 # :90-91
            lhs = E - viol
            two = (pr.g_L > -INF) & (pr.g_U < INF)
            lo = ~two & (pr.g_L > -INF)
            up = ~two & ~lo & (pr.g_U < INF)
            lhs = lhs + np.where(two, alpha * (ps[:, 0] - ps[:, 1]), 0.0)
            lhs = lhs + np.where(lo, alpha * ps[:, 0], 0.0)
            lhs = lhs - np.where(up, alpha * ps[:, 0], 0.0)
            phi += float(np.sum(self.nu * np.maximum(0.0, np.maximum(lhs - pr.g_U, pr.g_L - lhs))))
            return phi
        phi = pr.eval_f(xp)
        phi += float(np.sum(self.nu * self._row_viol(E)))
        return phi

    # slp.jl:122-147
    def compute_derivative(self):
        pr = self.problem
        if self.feasibility_restoration:
            D = float(np.sum(self.p_slack))
            viol = self._row_viol(self.E)
            lhs = self.E - viol
            D -= float(np.sum(self.nu * np.maximum(0.0, np.maximum(lhs - pr.g_U, pr.g_L - lhs))))
            return D
        D = float(self.df @ self.p)
        D -= float(np.sum(self.nu * self._row_viol(self.E)))
        return D

    def norm_violations(self, p=1):
        pr = self.problem
        return norm_violations(self.E, pr.g_L, pr.g_U, self.x, pr.x_L, pr.x_U, p)

    def kt_residuals(self):
        return kt_residuals(self.df, self.lam, self.mult_x_U, self.mult_x_L, self.jac())

    def norm_complementarity(self):
        pr = self.problem
        return norm_complementarity(self.E, pr.g_L, pr.g_U, self.lam)

    def finish(self):
        self.obj_val = self.problem.eval_f(self.x)
        self.status = int(self.ret)


class SlpLS(_Slp):
    """slp_line_search.jl."""

    def compute_nu(self):                                                 # :251-261
        if self.iter == 1:
            self.nu = np.abs(self.lam)
        else:
            self.nu = np.maximum(self.nu, np.abs(self.lam))

    def compute_alpha(self):                                              # :222-244
        o = self.options
        is_valid = True
        self.alpha = 1.0
        phi_x_p = self.compute_phi(self.x, self.alpha, self.p)
        while phi_x_p > self.phi + o.eta * self.alpha * self.directional_derivative:
            if self.alpha < o.min_alpha:
                if self.feasibility_restoration:
                    self.ret = -3
                is_valid = False
                break
            self.alpha *= o.tau
            phi_x_p = self.compute_phi(self.x, self.alpha, self.p)
        return is_valid

    def run(self):                                                        # :78-215
        o = self.options
        self.clip_start()
        self.iter = 1
        while True:
            self.eval_functions()
            self.alpha = 0.0
            self.prim_infeas = self.norm_violations(INF)
            self.dual_infeas = self.kt_residuals()
            self.compl = self.norm_complementarity()
            self.p, self.lam, self.mult_x_U, self.mult_x_L, self.p_slack, status = self.sub_optimize()
            if status not in (OPTIMAL, INFEASIBLE):
                self.ret = -3          # :129 compares (`slp.ret == -3`) where it means to assign; see DESIGN.md
                if self.prim_infeas <= o.tol_infeas:
                    self.ret = 6
                break
            elif status == INFEASIBLE:
                if self.feasibility_restoration:
                    self.ret = 6 if self.prim_infeas <= o.tol_infeas else 2
                    break
                self.feasibility_restoration = True
                continue
            self.compute_nu()
            self.phi = self.compute_phi(self.x, 0.0, self.p)
            self.directional_derivative = self.compute_derivative()
            is_valid_step = self.compute_alpha()
            if self.iter >= o.max_iter:
                self.ret = -1
                if self.prim_infeas <= o.tol_infeas:
                    self.ret = 6
                break
            if (self.prim_infeas <= o.tol_infeas and self.compl <= o.tol_residual) or \
                    float(np.max(np.abs(self.p))) <= o.tol_direction:
                if self.feasibility_restoration:
                    self.feasibility_restoration = False
                    self.iter += 1
                    continue
                elif self.dual_infeas <= o.tol_residual:
                    self.ret = 0
                    break
            if not is_valid_step:
                if self.ret == -3:
                    self.ret = 6 if self.prim_infeas <= o.tol_infeas else 2
                    break
                else:
                    self.feasibility_restoration = True
                self.iter += 1
                continue
            self.x = self.x + self.alpha * self.p
            self.iter += 1
        self.finish()
        return self


class SlpTR(_Slp):
    """slp_trust_region.jl."""

    def __init__(self, problem, options, engine_factory=None):
        super().__init__(problem, options, engine_factory)
        self.delta = options.tr_size
        self.delta_max = 2.0
        self.alpha1 = 0.1
        self.alpha2 = 0.25

    def compute_nu(self):                                                 # slp.jl:54-66
        if self.iter == 1:
            norm_df = 1.0 if self.feasibility_restoration else float(np.linalg.norm(self.df))
            rn = row_norms(self.jac())
            self.nu = np.maximum(1.0, norm_df / np.maximum(1.0, rn))
        else:
            self.nu = np.maximum(self.nu, np.abs(self.lam))

    def step_quality(self):                                               # :213-251
        o = self.options
        self.phi = self.compute_phi(self.x, 1.0, self.p) - self.compute_phi(self.x, 0.0, self.p)
        phi_pre = self.compute_derivative()
        if abs(phi_pre) > 0.0:
            rho = self.phi / phi_pre
            if rho <= 0:
                self.delta *= self.alpha1
            elif rho <= 0.25:
                self.delta *= self.alpha2
            elif rho > 0.75:
                self.delta = min(2 * self.delta, self.delta_max)
        else:
            rho = -self.phi
            if abs(self.phi) < 1.0e-8:
                if self.feasibility_restoration:
                    self.feasibility_restoration = False
                else:
                    if self.prim_infeas <= o.tol_infeas:
                        if self.dual_infeas <= o.tol_residual and self.compl <= o.tol_residual:
                            self.ret = 0
                        else:
                            self.ret = 6
                    else:
                        self.ret = 2
        return rho

    def run(self):                                                        # :87-206
        o = self.options
        pr = self.problem
        self.clip_start()
        self.iter = 1
        while True:
            self.eval_functions()
            self.p, self.lam, self.mult_x_U, self.mult_x_L, self.p_slack, status = self.sub_optimize(self.delta)
            if status not in (OPTIMAL, INFEASIBLE):
                # norm_violations(slp, slp.x): 1-norm, bound part on slp.x (App. C-8)
                self.ret = -3
                if norm_violations(pr.eval_g(self.x, np.zeros(pr.m)), pr.g_L, pr.g_U, self.x, pr.x_L, pr.x_U, 1) \
                        <= o.tol_infeas:
                    self.ret = 6
                break
            elif status == INFEASIBLE:
                if self.feasibility_restoration:
                    self.ret = 6 if self.prim_infeas <= o.tol_infeas else 2
                    break
                self.feasibility_restoration = True
                continue
            self.compute_nu()
            self.prim_infeas = self.norm_violations(INF)
            self.dual_infeas = self.kt_residuals()
            self.compl = self.norm_complementarity()
            if self.prim_infeas <= o.tol_infeas and self.compl <= o.tol_residual and \
                    float(np.max(np.abs(self.p))) <= o.tol_direction:
                if self.feasibility_restoration:
                    self.feasibility_restoration = False
                    self.iter += 1
                    if self.iter >= o.max_iter:          # guard: the reference loops forever here (see slp.py)
                        self.ret = 6 if self.prim_infeas <= o.tol_infeas else -1
                        break
                    continue
                elif self.dual_infeas <= o.tol_residual:
                    self.ret = 0
                    break
            if self.iter >= o.max_iter:
                self.ret = -1
                if self.prim_infeas <= o.tol_infeas:
                    self.ret = 6
                break
            rho = self.step_quality()
            if self.ret in (0, 2, 6):
                break
            if rho >= 0:
                self.x = self.x + self.p
            self.iter += 1
        self.finish()
        return self


def optimize(problem, options: Parameters, engine_factory=None):
    """model.jl:63-80 dispatch."""
    if options.algorithm == "Line Search":
        return SlpLS(problem, options, engine_factory).run()
    if options.algorithm == "Trust Region":
        return SlpTR(problem, options, engine_factory).run()
    raise ValueError("unknown algorithm")
