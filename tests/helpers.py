"""Shared test helpers: problem zoo and the oracle-side recording of sub-LP linearisations."""
import numpy as np

from activesetmethods_b200.examples import acopf, small_nlps
from oracle import slp_oracle as so

CASE3_PATH = "/root/reference/examples/acopf/case3.m"   # only read by the fixture generator, never by tests


def problem(name):
    if name == "toy":
        return small_nlps.ToyNlp()
    if name == "hs071":
        return small_nlps.Hs071()
    if name == "case9":
        return acopf.AcopfModel(acopf.case9())
    if name in acopf.PEGASE_SHAPES:
        return acopf.AcopfModel(acopf.synthetic_network(*acopf.PEGASE_SHAPES[name]))
    raise KeyError(name)


def record_sublps(pr, algorithm="Line Search", max_iter=40, limit=None):
    """Run the oracle SLP and return the list of sub-LP inputs it saw plus its per-LP log."""
    cls = so.SlpLS if algorithm == "Line Search" else so.SlpTR
    slp = cls(pr, so.Parameters(algorithm=algorithm, max_iter=max_iter))
    lps = []
    slp.record = lambda s, d: lps.append(d)
    slp.run()
    if limit is not None:
        lps = lps[:limit]
    return slp, lps


def lp_feasibility(K, x, lb, ub, rl, ru):
    Kx = K @ x
    row = max(float(np.max(np.maximum(0.0, rl - Kx), initial=0.0)), float(np.max(np.maximum(0.0, Kx - ru), initial=0.0)))
    col = max(float(np.max(np.maximum(0.0, lb - x), initial=0.0)), float(np.max(np.maximum(0.0, x - ub), initial=0.0)))
    return max(row, col)


def dual_feasibility(K, cost, x, y, lb, ub, tol=1e-9):
    """Infinity norm of the part of the reduced cost c - K'y that no active bound explains."""
    rc = cost - K.T @ y
    at_l = x <= lb + tol
    at_u = x >= ub - tol
    bad = np.where(at_l & at_u, 0.0, np.where(at_l, np.minimum(rc, 0.0), np.where(at_u, np.maximum(rc, 0.0), rc)))
    return float(np.max(np.abs(bad), initial=0.0))
