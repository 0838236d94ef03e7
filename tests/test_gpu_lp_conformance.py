"""the generic LP handle (asm_lp_*, what a GLPK-replacing MOI optimizer binds) on the LP subset of
MathOptInterface's ``contlinear`` suite the reference runs through GLPK (reference test/MOI_wrapper.jl:72-88): the
situations those cases exercise are restated here as explicit small LPs with known answers, each cross-checked with
HiGHS -- infeasible, unbounded but boxed (the reference always boxes its sub-LP, subproblem.jl:427-434), equality
only, free column inside a box, free and empty rows, duplicate and zero coefficients, maximisation through the
reference's sign flip (MOI_wrapper.jl:1037-1054), degenerate optimum, fixed columns, range rows, dual signs
(GreaterThan >= 0, LessThan <= 0: the convention consumed at src/algorithms/common.jl:38).  Run on both engine
families (0: barrier engine, 5: PDHG)."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import lp_feasibility
from oracle import slp_oracle as so

pytestmark = pytest.mark.gpu
INF = np.inf
ENGINES = [0, 5]


def _solve(K, c, lb, ub, rl, ru, engine, c0=0.0, **kw):
    from activesetmethods_b200.sublp import B200LP
    K = sp.csr_matrix(K)
    lp = B200LP(K.shape[1], K.shape[0], K.indptr, K.indices, engine=engine, eps_rel=1e-8, **kw)
    lp.set_matrix_values(K.data)
    lp.set_objective(np.asarray(c, float), c0)
    lp.set_col_bounds(np.asarray(lb, float), np.asarray(ub, float))
    lp.set_row_bounds(np.asarray(rl, float), np.asarray(ru, float))
    info = lp.optimize()[0]
    out = dict(info=info, x=lp.primal().copy(), y=lp.row_dual().copy(), z=lp.col_dual())
    lp.close()
    return out


def _highs(K, c, lb, ub, rl, ru, c0=0.0):
    st, x, y, z, obj = so.HighsLp().solve(sp.csc_matrix(K), np.asarray(c, float), c0, np.asarray(lb, float),
                                          np.asarray(ub, float), np.asarray(rl, float), np.asarray(ru, float))
    return st, x, y, z, obj


@pytest.mark.parametrize("engine", ENGINES)
def test_linear1_like(gpu, engine):
    """max x + z ... restated as min: min -x  s.t.  x + y <= 1, x, y >= 0  ->  x = 1, objective -1, row dual -1."""
    K = np.array([[1.0, 1.0]])
    out = _solve(K, [-1.0, 0.0], [0, 0], [INF, INF], [-INF], [1.0], engine)
    assert out["info"]["status"] == 0
    assert abs(out["info"]["objective"] + 1.0) <= 1e-7
    assert np.allclose(out["x"], [1.0, 0.0], atol=1e-6)
    assert out["y"][0] <= 1e-9 and abs(out["y"][0] + 1.0) <= 1e-6           # LessThan row: dual <= 0
    lo, up = out["z"]
    assert np.all(lo >= -1e-12) and np.all(up <= 1e-12)                      # bound duals: >= 0 lower, <= 0 upper


@pytest.mark.parametrize("engine", ENGINES)
def test_infeasible_rows_against_bounds(gpu, engine):
    """x >= 1 as a row, x <= 0 as a bound: INFEASIBLE (the status SLP branches on, slp_line_search.jl:134-147)."""
    out = _solve(np.array([[1.0]]), [1.0], [-5.0], [0.0], [1.0], [INF], engine)
    assert out["info"]["status"] == 1
    st = _highs(np.array([[1.0]]), [1.0], [-5.0], [0.0], [1.0], [INF])[0]
    assert st == so.INFEASIBLE


@pytest.mark.parametrize("engine", ENGINES)
def test_infeasible_pair_of_rows(gpu, engine):
    """x + y >= 2 and x + y <= 1 inside a box that would allow either."""
    K = np.array([[1.0, 1.0], [1.0, 1.0]])
    out = _solve(K, [1.0, 1.0], [-10, -10], [10, 10], [2.0, -INF], [INF, 1.0], engine)
    assert out["info"]["status"] == 1


@pytest.mark.parametrize("engine", ENGINES)
def test_unbounded_direction_stopped_by_the_box(gpu, engine):
    """min -x - 2y over the +-1000 box of the reference's line search (slp.jl:23) with one inactive row."""
    K = np.array([[1.0, -1.0]])
    out = _solve(K, [-1.0, -2.0], [-1000, -1000], [1000, 1000], [-5000.0], [5000.0], engine)
    assert out["info"]["status"] == 0
    assert np.allclose(out["x"], [1000.0, 1000.0], rtol=1e-7)
    assert abs(out["info"]["objective"] + 3000.0) <= 1e-6 * 3000


@pytest.mark.parametrize("engine", ENGINES)
def test_equality_only_with_free_column_inside_box(gpu, engine):
    """Two equalities, three columns; the third column's cost is zero and it ends strictly inside its box."""
    K = np.array([[1.0, 1.0, 0.0], [1.0, -1.0, 1.0]])
    c = [1.0, 2.0, 0.0]
    lb, ub = [-3, -3, -100], [3, 3, 100]
    rl = ru = [1.0, 0.5]
    out = _solve(K, c, lb, ub, rl, ru, engine)
    st, x, y, z, obj = _highs(K, c, lb, ub, rl, ru)
    assert out["info"]["status"] == st == 0
    assert abs(out["info"]["objective"] - obj) <= 1e-7 * max(1.0, abs(obj))
    assert lp_feasibility(sp.csr_matrix(K), out["x"], np.array(lb, float), np.array(ub, float), np.array(rl), np.array(ru)) <= 1e-7
    lo, up = out["z"]
    assert abs(lo[2]) <= 1e-7 and abs(up[2]) <= 1e-7                        # free inside the box: no bound dual


@pytest.mark.parametrize("engine", ENGINES)
def test_free_and_empty_rows(gpu, engine):
    """Row 1 has no finite bound, row 2 has no entries but bounds that contain 0: neither changes the optimum."""
    K = sp.csr_matrix((np.array([1.0, 1.0, 2.0, -1.0]), np.array([0, 1, 0, 1]), np.array([0, 2, 4, 4])), shape=(3, 2))
    c = [-1.0, -1.0]
    out = _solve(K, c, [0, 0], [4, 4], [-INF, -INF, -1.0], [3.0, INF, 1.0], engine)
    st, x, y, z, obj = _highs(K, c, [0, 0], [4, 4], [-INF, -INF, -1.0], [3.0, INF, 1.0])
    assert out["info"]["status"] == st == 0
    assert abs(out["info"]["objective"] - obj) <= 1e-7 * max(1.0, abs(obj))
    assert abs(out["y"][1]) <= 1e-9 and abs(out["y"][2]) <= 1e-9


@pytest.mark.parametrize("engine", ENGINES)
def test_empty_row_with_impossible_bounds_is_infeasible(gpu, engine):
    K = sp.csr_matrix((np.array([1.0]), np.array([0]), np.array([0, 1, 1])), shape=(2, 1))
    out = _solve(K, [1.0], [0.0], [1.0], [-INF, 1.0], [5.0, 2.0], engine)       # 0 in [1, 2] is impossible
    assert out["info"]["status"] == 1


def test_duplicate_coefficients_are_summed(gpu):
    """The same (row, column) twice in the pattern, as MOI allows in a ScalarAffineFunction: the barrier engine needs
    a deduplicated CSR and refuses the pattern, so engine 0 falls back to PDHG; both duplicates count."""
    K = sp.csr_matrix((np.array([1.0, 1.0, 1.0]), np.array([0, 0, 1]), np.array([0, 3])), shape=(1, 2))   # 2x + y <= 2
    for engine in (0, 5):
        out = _solve(K, [-1.0, -0.25], [0, 0], [10, 10], [-INF], [2.0], engine)
        assert out["info"]["status"] == 0
        assert abs(out["info"]["objective"] + 0.5 * 2.0) <= 1e-6               # x = 1 (2x = 2), y = 0 -> -1


@pytest.mark.parametrize("engine", ENGINES)
def test_explicit_zero_coefficients(gpu, engine):
    K = sp.csr_matrix((np.array([1.0, 0.0, 0.0, 1.0]), np.array([0, 1, 0, 1]), np.array([0, 2, 4])), shape=(2, 2))
    c = [1.0, -1.0]
    out = _solve(K, c, [-2, -2], [2, 2], [0.5, -INF], [INF, 1.5], engine)
    assert out["info"]["status"] == 0
    assert np.allclose(out["x"], [0.5, 1.5], atol=1e-6)


@pytest.mark.parametrize("engine", ENGINES)
def test_max_sense_through_the_sign_flip(gpu, engine):
    """MOI_wrapper.jl:1037-1054 turns MAX_SENSE into min (-f): the maximiser is the minimiser of the flipped costs."""
    rng = np.random.default_rng(3)
    K = sp.random(6, 8, density=0.5, random_state=3, format="csr")
    K.data[:] = rng.standard_normal(K.nnz)
    c = rng.standard_normal(8)
    lb, ub = np.full(8, -1.0), np.full(8, 2.0)
    x0 = rng.uniform(-0.5, 1.5, 8)
    rl, ru = K @ x0 - 0.3, K @ x0 + 0.3
    mn = _solve(K, c, lb, ub, rl, ru, engine)
    mx = _solve(K, -c, lb, ub, rl, ru, engine)
    r_mn = _highs(K, c, lb, ub, rl, ru)[4]
    r_mx = -_highs(K, -c, lb, ub, rl, ru)[4]
    assert mn["info"]["status"] == mx["info"]["status"] == 0
    assert abs(mn["info"]["objective"] - r_mn) <= 1e-7 * max(1.0, abs(r_mn))
    assert abs(-mx["info"]["objective"] - r_mx) <= 1e-7 * max(1.0, abs(r_mx))
    assert r_mx >= r_mn - 1e-9


@pytest.mark.parametrize("engine", ENGINES)
def test_degenerate_optimum_and_fixed_columns(gpu, engine):
    """A face of minimisers (cost parallel to a constraint) and a column with lb == ub."""
    K = np.array([[1.0, 1.0, 1.0], [1.0, -1.0, 0.0]])
    c = [1.0, 1.0, 0.0]
    lb, ub = [0, 0, 0.25], [5, 5, 0.25]
    out = _solve(K, c, lb, ub, [2.0, -1.0], [INF, 1.0], engine)
    st, x, y, z, obj = _highs(K, c, lb, ub, [2.0, -1.0], [INF, 1.0])
    assert out["info"]["status"] == st == 0
    assert abs(out["info"]["objective"] - obj) <= 1e-7 * max(1.0, abs(obj))
    assert out["x"][2] == 0.25


@pytest.mark.parametrize("engine", ENGINES)
def test_range_rows_and_dual_signs_against_highs(gpu, engine):
    rng = np.random.default_rng(11)
    n, m = 12, 9
    K = sp.random(m, n, density=0.45, random_state=11, format="csr")
    K.data[:] = rng.standard_normal(K.nnz)
    c = rng.standard_normal(n)
    x0 = rng.uniform(-1, 1, n)
    rl, ru = K @ x0 - rng.uniform(0.1, 1, m), K @ x0 + rng.uniform(0.1, 1, m)
    rl[:3] = ru[:3]
    lb, ub = np.full(n, -1.5), np.full(n, 1.5)
    out = _solve(K, c, lb, ub, rl, ru, engine)
    st, x, y, z, obj = _highs(K, c, lb, ub, rl, ru)
    assert out["info"]["status"] == st == 0
    assert abs(out["info"]["objective"] - obj) <= 1e-7 * max(1.0, abs(obj))
    assert lp_feasibility(K, out["x"], lb, ub, rl, ru) <= 1e-7
    Kx = K @ out["x"]
    yv = out["y"]
    tol = 1e-6
    assert np.all(yv[(Kx > rl + tol) & (Kx < ru - tol)] <= tol) and np.all(yv[(Kx > rl + tol) & (Kx < ru - tol)] >= -tol)
    assert np.all(yv[(Kx <= rl + tol) & (rl < ru)] >= -tol)                  # on the lower side: dual >= 0
    assert np.all(yv[(Kx >= ru - tol) & (rl < ru)] <= tol)                   # on the upper side: dual <= 0
    lo, up = out["z"]
    rc = c - K.T @ yv
    assert np.linalg.norm(rc - lo - up) <= 1e-6 * (1.0 + np.linalg.norm(c))   # c - K'y - z_L - z_U = 0
    if engine == 0:   # unique optimum: the vertex solver's duals are the barrier engine's
        assert np.allclose(yv, y, atol=1e-5)
