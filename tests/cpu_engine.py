"""A CPU stand-in with the interface of ``activesetmethods_b200.sublp.SubLp`` (batch 1), built on the oracle.
Test infrastructure only: it lets the host drivers of ``activesetmethods_b200/slp.py`` run without a GPU, so their
control flow can be compared with the oracle's independent restatement of the reference drivers (same LP solver
underneath, hence the same trajectory)."""
import numpy as np

from oracle import slp_oracle as so


class OracleEngine:
    def __init__(self, n, m, j_str, x_L, x_U, g_L, g_U, **_):
        self.n, self.m = n, m
        self.x_L, self.x_U, self.g_L, self.g_U = (np.asarray(a, float) for a in (x_L, x_U, g_L, g_U))
        self.pat = so.JacobianPattern(m, n, j_str)
        self.lp = so.SubLp(self.pat, self.g_L, self.g_U, self.x_L, self.x_U)
        self.last_info = None
        self._d = None
        self._slack = np.zeros((m, 2))
        self._p = np.zeros(n)

    # -- hot path
    def update(self, x_k, f, df, E, dE, delta, feasibility=False):
        self._d = dict(x=np.array(x_k, float), f=float(f), df=np.array(df, float), E=np.array(E, float),
                       dE=np.array(dE, float), delta=float(delta), fr=bool(feasibility))
        self._vals = self.pat.assemble(self._d["dE"])
        self._J = self.pat.matrix(self._vals)

    def solve_extract(self):
        d = self._d
        p, lam, mu_u, mu_l, slack, status = self.lp.solve(self._vals, d["df"], d["f"], d["E"], d["x"], d["delta"], d["fr"])
        self.last_info = [dict(status=status, objective=self.lp.last_objective, iterations=0)]
        self._p = p
        self._slack = slack if slack is not None else np.zeros((self.m, 2))
        return p, lam, mu_u, mu_l, self._slack, status

    def sub_optimize(self, x_k, f, df, E, dE, delta=1000.0, feasibility=False):
        self.update(x_k, f, df, E, dE, delta, feasibility)
        return self.solve_extract()

    # -- reductions (same conventions as SubLp: None = the data of the last update)
    def norm_violations(self, E=None, x=None, p=1):
        E = self._d["E"] if E is None else np.asarray(E, float)
        x = self._d["x"] if x is None else np.asarray(x, float)
        return so.norm_violations(E, self.g_L, self.g_U, x, self.x_L, self.x_U, p)

    def kt_residuals(self, lam, mult_x_U, mult_x_L, df=None):
        return so.kt_residuals(self._d["df"] if df is None else df, lam, mult_x_U, mult_x_L, self._J)

    def norm_complementarity(self, lam, E=None):
        return so.norm_complementarity(self._d["E"] if E is None else E, self.g_L, self.g_U, lam)

    def row_norms(self):
        return so.row_norms(self._J)

    def _viol(self, E):
        return np.maximum(0.0, np.maximum(E - self.g_U, self.g_L - E))

    def merit_phi(self, base, E_trial, nu, alpha, feasibility=False):
        E = self._d["E"] if E_trial is None else np.asarray(E_trial, float)
        if not feasibility:
            return float(base) + float(np.sum(nu * self._viol(E)))
        ps = self._slack
        two = (self.g_L > -np.inf) & (self.g_U < np.inf)
        lo = ~two & (self.g_L > -np.inf)
        up = ~two & ~lo & (self.g_U < np.inf)
        lhs = E - self._viol(self._d["E"])
        lhs = lhs + np.where(two, alpha * (ps[:, 0] - ps[:, 1]), 0.0) + np.where(lo, alpha * ps[:, 0], 0.0) \
            - np.where(up, alpha * ps[:, 0], 0.0)
        return float(base) + alpha * float(np.sum(ps)) + \
            float(np.sum(nu * np.maximum(0.0, np.maximum(lhs - self.g_U, self.g_L - lhs))))

    def merit_derivative(self, nu, feasibility=False):
        E = self._d["E"]
        if feasibility:
            lhs = E - self._viol(E)
            return float(np.sum(self._slack)) - \
                float(np.sum(nu * np.maximum(0.0, np.maximum(lhs - self.g_U, self.g_L - lhs))))
        return float(self._d["df"] @ self._p) - float(np.sum(nu * self._viol(E)))

    def close(self):
        pass


class OracleBatchEngine:
    """Batch twin of ``OracleEngine`` (one oracle engine per scenario) with the ``SubLp(batch=B)`` conventions:
    arrays are ``[B, ...]``, the restoration flag is shared by the call."""

    def __init__(self, n, m, j_str, x_L, x_U, g_L, g_U, batch=1, **_):
        self.B = batch
        self.eng = [OracleEngine(n, m, j_str, x_L[s], x_U[s], g_L[s], g_U[s]) for s in range(batch)]
        self.last_info = None

    def update(self, x_k, f, df, E, dE, delta, feasibility=False):
        for s, e in enumerate(self.eng):
            e.update(x_k[s], f[s], df[s], E[s], dE[s], delta, feasibility)

    def solve_extract(self):
        outs = [e.solve_extract() for e in self.eng]
        self.last_info = [e.last_info[0] for e in self.eng]
        return tuple(np.array([o[k] for o in outs]) for k in range(6))

    def norm_violations(self, E=None, x=None, p=1):
        return np.array([e.norm_violations(None if E is None else E[s], None if x is None else x[s], p)
                         for s, e in enumerate(self.eng)])

    def kt_residuals(self, lam, mu_u, mu_l):
        return np.array([e.kt_residuals(lam[s], mu_u[s], mu_l[s]) for s, e in enumerate(self.eng)])

    def norm_complementarity(self, lam):
        return np.array([e.norm_complementarity(lam[s]) for s, e in enumerate(self.eng)])

    def merit_phi(self, base, E_trial, nu, alpha, feasibility=False):
        return np.array([e.merit_phi(base[s], None if E_trial is None else E_trial[s], nu[s], alpha[s], feasibility)
                         for s, e in enumerate(self.eng)])

    def merit_derivative(self, nu, feasibility=False):
        return np.array([e.merit_derivative(nu[s], feasibility) for s, e in enumerate(self.eng)])
