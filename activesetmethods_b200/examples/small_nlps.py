"""Small host-side NLPs used by the reference's own tests, restated as callback bundles.

* ``ToyNlp``   – reference ``test/ext_solver.jl:13-19`` / ``examples/toy_example.jl:13-19``:
                 min X^2 + X  s.t.  X^2 - X == 2,  X*Y == 1,  X*Y >= 0,  X >= -2   → X = Y = -1
                 (``test/runtests.jl:11-13``).  ``X >= -2`` is a *linear row* (JuMP ``@constraint``), so the
                 row order of ``src/MOI_wrapper.jl:683-689`` is [lin>= row, NLP rows...].
* ``Hs071``    – Hock–Schittkowski 71, exercised through ``MOIT.nlptest`` (``test/MOI_wrapper.jl:109``);
                 optimum 17.0140173.
* ``RandomNlp``– dense-ish random quadratic-constraint problem with *duplicate* COO entries, to exercise the
                 ordered duplicate summation of ``src/algorithms/common.jl:12-20``.

Every class exposes the tuple the reference's ``Model`` constructor takes (``src/model.jl:33-60``):
``n, m, x_L, x_U, g_L, g_U, j_str`` (1-based (row, col) pairs), ``x0`` and the callbacks
``eval_f(x)``, ``eval_grad_f(x, grad)``, ``eval_g(x, g)``, ``eval_jac_g(x, mode, rows, cols, values)``.
"""
from __future__ import annotations

import numpy as np

INF = np.inf


class _Base:
    def eval_jac_structure(self, rows, cols):
        rows[:] = self.j_str[:, 0]
        cols[:] = self.j_str[:, 1]


class ToyNlp(_Base):
    n, m = 2, 4

    def __init__(self):
        self.x_L = np.array([-INF, -INF])
        self.x_U = np.array([INF, INF])
        self.g_L = np.array([-2.0, 2.0, 1.0, 0.0])
        self.g_U = np.array([INF, 2.0, 1.0, INF])
        self.j_str = np.array([[1, 1], [2, 1], [3, 1], [3, 2], [4, 1], [4, 2]], dtype=np.int64)
        self.nnz = 6
        # no start given, no bounds: min(0, +Inf) = 0  (src/MOI_wrapper.jl:1113-1130)
        self.x0 = np.zeros(2)

    def eval_f(self, x):
        return float(x[0] * x[0] + x[0])

    def eval_grad_f(self, x, g):
        g[0] = 2.0 * x[0] + 1.0
        g[1] = 0.0
        return g

    def eval_g(self, x, g):
        g[0] = x[0]
        g[1] = x[0] * x[0] - x[0]
        g[2] = x[0] * x[1]
        g[3] = x[0] * x[1]
        return g

    def eval_jac_g(self, x, mode, rows, cols, values):
        if mode == "Structure":
            return self.eval_jac_structure(rows, cols)
        values[0] = 1.0
        values[1] = 2.0 * x[0] - 1.0
        values[2] = x[1]
        values[3] = x[0]
        values[4] = x[1]
        values[5] = x[0]
        return values


class Hs071(_Base):
    n, m = 4, 2

    def __init__(self):
        self.x_L = np.full(4, 1.0)
        self.x_U = np.full(4, 5.0)
        self.g_L = np.array([25.0, 40.0])
        self.g_U = np.array([INF, 40.0])
        self.j_str = np.array([[1, 1], [1, 2], [1, 3], [1, 4], [2, 1], [2, 2], [2, 3], [2, 4]], dtype=np.int64)
        self.nnz = 8
        self.x0 = np.array([1.0, 5.0, 5.0, 1.0])

    def eval_f(self, x):
        return float(x[0] * x[3] * (x[0] + x[1] + x[2]) + x[2])

    def eval_grad_f(self, x, g):
        g[0] = x[3] * (2 * x[0] + x[1] + x[2])
        g[1] = x[0] * x[3]
        g[2] = x[0] * x[3] + 1.0
        g[3] = x[0] * (x[0] + x[1] + x[2])
        return g

    def eval_g(self, x, g):
        g[0] = x[0] * x[1] * x[2] * x[3]
        g[1] = float(np.dot(x, x))
        return g

    def eval_jac_g(self, x, mode, rows, cols, values):
        if mode == "Structure":
            return self.eval_jac_structure(rows, cols)
        values[0] = x[1] * x[2] * x[3]
        values[1] = x[0] * x[2] * x[3]
        values[2] = x[0] * x[1] * x[3]
        values[3] = x[0] * x[1] * x[2]
        values[4:8] = 2.0 * x
        return values


class RandomNlp(_Base):
    """min c'x + 0.5 sum d_j x_j^2  s.t.  rows  a_i'x + 0.5 x'Q_i x (diagonal Q_i) in [gL, gU], box.
    The Jacobian pattern lists the affine and the quadratic contribution of the same (row, col) as two
    separate COO entries — duplicates, as ``append_to_jacobian_sparsity!`` produces for a
    ``ScalarQuadraticFunction`` (``src/MOI_wrapper.jl:699-713``)."""

    def __init__(self, n=12, m=9, density=0.4, seed=0, n_eq=3, n_range=2):
        rng = np.random.default_rng(seed)
        self.n, self.m = n, m
        mask = rng.random((m, n)) < density
        mask[np.arange(m), rng.integers(0, n, m)] = True
        self.A = np.where(mask, rng.standard_normal((m, n)), 0.0)
        self.Q = np.where(mask & (rng.random((m, n)) < 0.5), 0.3 * rng.standard_normal((m, n)), 0.0)
        self.c = rng.standard_normal(n)
        self.d = rng.uniform(0.1, 1.0, n)
        xs = rng.uniform(-0.5, 0.5, n)                       # a feasible point
        gs = self.A @ xs + 0.5 * self.Q @ (xs * xs)
        gl = np.full(m, -INF)
        gu = np.full(m, INF)
        kind = np.array(["eq"] * n_eq + ["rng"] * n_range + ["lo", "up"] * m)[:m]
        for i, k in enumerate(kind):
            if k == "eq":
                gl[i] = gu[i] = gs[i]
            elif k == "rng":
                gl[i], gu[i] = gs[i] - 0.3, gs[i] + 0.4
            elif k == "lo":
                gl[i] = gs[i] - 0.2
            else:
                gu[i] = gs[i] + 0.2
        self.g_L, self.g_U = gl, gu
        self.x_L = np.full(n, -2.0)
        self.x_U = np.full(n, 2.0)
        self.x_L[::5] = -INF
        self.x_U[1::7] = INF
        rows, cols, kinds = [], [], []
        for i in range(m):
            for j in np.nonzero(mask[i])[0]:
                rows.append(i + 1); cols.append(j + 1); kinds.append(0)
            for j in np.nonzero(self.Q[i])[0]:
                rows.append(i + 1); cols.append(j + 1); kinds.append(1)
        self.j_str = np.stack([rows, cols], 1).astype(np.int64)
        self._kind = np.array(kinds)
        self.nnz = len(rows)
        self.x0 = np.zeros(n)

    def eval_f(self, x):
        return float(self.c @ x + 0.5 * np.sum(self.d * x * x))

    def eval_grad_f(self, x, g):
        g[:] = self.c + self.d * x
        return g

    def eval_g(self, x, g):
        g[:] = self.A @ x + 0.5 * self.Q @ (x * x)
        return g

    def eval_jac_g(self, x, mode, rows, cols, values):
        if mode == "Structure":
            return self.eval_jac_structure(rows, cols)
        r = self.j_str[:, 0] - 1
        c = self.j_str[:, 1] - 1
        values[:] = np.where(self._kind == 0, self.A[r, c], self.Q[r, c] * x[c])
        return values
