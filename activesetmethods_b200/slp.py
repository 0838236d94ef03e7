"""SLP drivers on top of the GPU sub-LP engine — the host orchestration of the reference, unchanged in
behaviour, with every array-sized operation of the iteration moved behind the C ABI.

================================  ==========================================================================
here                              reference (relative to /root/reference)
================================  ==========================================================================
``Parameters``                    ``src/parameters.jl:1-29``
``Model`` / ``optimize``          ``src/model.jl:1-80`` (problem container, algorithm dispatch)
``SlpLS``                         ``src/algorithms/slp_line_search.jl:4-261``
``SlpTR``                         ``src/algorithms/slp_trust_region.jl:10-251``, ``src/algorithms/slp.jl:54-66``
``STATUS``                        ``src/status.jl:2-22``
================================  ==========================================================================

The NLP callbacks (``eval_f``, ``eval_grad_f``, ``eval_g``, ``eval_jac_g``) stay on the host exactly like the
JuMP ``NLPEvaluator`` of the reference (``src/MOI_wrapper.jl:1047-1069``).  Per iteration the driver hands
``x, f, df, E, dE`` to ``SubLp.sub_optimize`` (one H2D hand-over, assembly + bounds + PDHG on the device,
one D2H read-back) and asks the device for the merit / KKT reductions.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

import numpy as np

from .sublp import SubLp, LP_OPTIMAL, LP_INFEASIBLE

INF = math.inf

# src/status.jl:2-22
STATUS = {
    0: "Solve_Succeeded", 1: "Solved_To_Acceptable_Level", 2: "Infeasible_Problem_Detected",
    3: "Search_Direction_Becomes_Too_Small", 4: "Diverging_Iterates", 5: "User_Requested_Stop",
    6: "Feasible_Point_Found", -1: "Maximum_Iterations_Exceeded", -2: "Restoration_Failed",
    -3: "Error_In_Step_Computation", -4: "Maximum_CpuTime_Exceeded", -5: "Optimize_not_called",
    -10: "Not_Enough_Degrees_Of_Freedom", -11: "Invalid_Problem_Definition", -12: "Invalid_Option",
    -13: "Invalid_Number_Detected", -100: "Unrecoverable_Exception", -101: "NonIpopt_Exception_Thrown",
    -102: "Insufficient_Memory", -199: "Internal_Error",
}


@dataclass
class Parameters:
    """src/parameters.jl:1-29.  ``external_optimizer`` is the hook: a callable
    ``(n, m, j_str, x_L, x_U, g_L, g_U) -> SubLp``-like object, or ``None`` → status -12 (model.jl:64-66).
    The default is the B200 engine."""
    mu_merit: float = 1.0
    external_optimizer: object = "B200LP"
    method: str = "SLP"
    algorithm: str = "Line Search"
    max_mu: float = 1.0e10
    rho: float = 0.8
    eta: float = 0.4
    tau: float = 0.9
    tol_direction: float = 1.0e-6
    tol_residual: float = 0.01
    tol_infeas: float = 0.01
    max_iter: int = 1000
    time_limit: float = INF
    min_alpha: float = 1.0e-6
    tr_size: float = 0.4
    OutputFlag: int = 0
    StatisticsFlag: int = 0
    # engine options forwarded to asm_lp_params (not in the reference)
    lp_options: dict = field(default_factory=dict)
    device: int = 0
    # line search on an examples.acopf.AcopfModel with the B200 engine: evaluate f, g, the Jacobian and the
    # backtracking trials with the device-side ACOPF evaluator instead of the host callbacks
    device_evaluator: bool = False


class Model:
    """src/model.jl:1-61: the problem container handed over by the MOI front-end
    (``src/MOI_wrapper.jl:1093-1099``)."""

    def __init__(self, n, m, x_L, x_U, g_L, g_U, j_str, eval_f, eval_g, eval_grad_f, eval_jac_g,
                 parameters: Parameters | None = None, x0=None):
        self.n, self.m = int(n), int(m)
        self.x_L = np.asarray(x_L, dtype=float)
        self.x_U = np.asarray(x_U, dtype=float)
        self.g_L = np.asarray(g_L, dtype=float)
        self.g_U = np.asarray(g_U, dtype=float)
        self.j_str = np.asarray(j_str, dtype=np.int64).reshape(-1, 2)
        self.eval_f, self.eval_g, self.eval_grad_f, self.eval_jac_g = eval_f, eval_g, eval_grad_f, eval_jac_g
        self.parameters = parameters if parameters is not None else Parameters()
        self.x = np.zeros(self.n) if x0 is None else np.array(x0, dtype=float)
        self.g = np.zeros(self.m)
        self.mult_g = np.zeros(self.m)
        self.mult_x_L = np.zeros(self.n)
        self.mult_x_U = np.zeros(self.n)
        self.obj_val = 0.0
        self.status = -5
        self.statistics = {}

    @classmethod
    def from_problem(cls, problem, parameters=None):
        """Wrap an object exposing ``n, m, x_L, x_U, g_L, g_U, j_str, x0`` and the four callbacks."""
        mdl = cls(problem.n, problem.m, problem.x_L, problem.x_U, problem.g_L, problem.g_U, problem.j_str,
                  problem.eval_f, problem.eval_g, problem.eval_grad_f, problem.eval_jac_g, parameters, problem.x0)
        mdl.source = problem         # the device-side evaluator needs the network behind the callbacks
        return mdl

    def add_statistic(self, name, value):                                   # model.jl:82-97
        if self.parameters.StatisticsFlag:
            self.statistics.setdefault(name, []).append(value)


def optimize(model: Model):
    """model.jl:63-80."""
    o = model.parameters
    if o.external_optimizer is None:
        model.status = -12
        return model
    if o.method == "SLP" and o.algorithm == "Line Search":
        SlpLS(model).run()
    elif o.method == "SLP" and o.algorithm == "Trust Region":
        SlpTR(model).run()
    else:
        raise ValueError(f"unknown method/algorithm {o.method}/{o.algorithm}")
    return model


class _Slp:
    def __init__(self, model: Model):
        self.problem = model
        self.options = model.parameters
        n, m = model.n, model.m
        self.x = model.x.copy()
        self.p = np.zeros(n)
        self.p_slack = np.zeros((m, 2))
        self.lam = np.zeros(m)
        self.mult_x_L = np.zeros(n)
        self.mult_x_U = np.zeros(n)
        self.f = 0.0
        self.df = np.zeros(n)
        self.E = np.zeros(m)
        self.dE = np.zeros(len(model.j_str))
        self.phi = INF
        self.nu = np.zeros(m)
        self.alpha = 1.0
        self.directional_derivative = 0.0
        self.prim_infeas = INF
        self.dual_infeas = INF
        self.compl = INF
        self.feasibility_restoration = False
        self.iter = 1
        self.ret = -5
        self.optimizer = None
        self.lp_log = []        # (status, objective, restoration flag, PDHG iterations) per sub-LP
        self.lp_time = 0.0
        self.start_time = 0.0
        self.record = None      # optional callable(slp, dict) invoked before every sub-LP (tests / benchmarks)

    # slp.jl:186-191
    def eval_functions(self):
        pr = self.problem
        self.f = pr.eval_f(self.x)
        pr.eval_grad_f(self.x, self.df)
        pr.eval_g(self.x, self.E)
        pr.eval_jac_g(self.x, "eval", None, None, self.dE)

    def _instantiate(self):
        """slp.jl:24-37: lazy creation of the external optimizer with the problem skeleton."""
        pr, o = self.problem, self.options
        factory = o.external_optimizer
        if factory == "B200LP":
            # Line search: every sub-LP starts from the previous one's (p, lambda) -- the analogue of GLPK keeping
            # its basis (one glp_prob per SLP run, slp.jl:24); it halves the PDHG work (case118: 1.03 M -> 0.51 M
            # iterations for the same 14 SLP iterations).  Trust region: cold starts; there a warm start changes
            # which minimiser of the (degenerate) restoration LPs is returned and the runs measured so far ended
            # further from the optimum.  ``lp_options={"warm_start": ...}`` overrides either.
            warm = 1 if o.algorithm == "Line Search" else 0
            return SubLp(pr.n, pr.m, pr.j_str, pr.x_L, pr.x_U, pr.g_L, pr.g_U, batch=1, device=o.device,
                         **{"warm_start": warm, **o.lp_options})
        return factory(pr.n, pr.m, pr.j_str, pr.x_L, pr.x_U, pr.g_L, pr.g_U)

    # slp.jl:23-47
    def sub_optimize(self, delta=1000.0, updated=False):
        """``updated``: the data push for this iterate has already been done by ``optimizer.update``."""
        if self.optimizer is None:
            self.optimizer = self._instantiate()
        if self.record is not None:
            self.record(self, dict(x=self.x.copy(), f=self.f, df=self.df.copy(), E=self.E.copy(),
                                   dE=self.dE.copy(), delta=delta, fr=self.feasibility_restoration))
        t0 = time.time()
        if updated:
            out = self.optimizer.solve_extract()
        else:
            out = self.optimizer.sub_optimize(self.x, self.f, self.df, self.E, self.dE, delta,
                                              self.feasibility_restoration)
        self.lp_time = time.time() - t0
        info = self.optimizer.last_info[0] if getattr(self.optimizer, "last_info", None) else {}
        self.lp_log.append((out[5], info.get("objective"), self.feasibility_restoration, info.get("iterations")))
        return out

    def clip_start(self):
        pr = self.problem                                                   # slp_line_search.jl:96-105
        lo = pr.x_L > -INF
        self.x[lo] = np.maximum(self.x[lo], pr.x_L[lo])
        up = pr.x_U > -INF                                                  # sic
        self.x[up] = np.minimum(self.x[up], pr.x_U[up])

    # slp.jl:79-115 — f / g at the trial point on the host (NLP evaluator), the m-length reduction on the GPU
    def compute_phi(self, x, alpha, p):
        pr = self.problem
        if self.options.device_evaluator and alpha != 0.0:      # f and g at the trial point on the device
            return self.optimizer.acopf_trial(alpha, self.nu, self.prim_infeas if self.feasibility_restoration else None,
                                              self.feasibility_restoration)
        xp = x + alpha * p
        E = None if alpha == 0.0 else pr.eval_g(xp, np.zeros(pr.m))
        if self.feasibility_restoration:
            return self.optimizer.merit_phi(self.prim_infeas, E, self.nu, alpha, True)
        base = self.f if (self.options.device_evaluator and alpha == 0.0) else pr.eval_f(xp)
        return self.optimizer.merit_phi(base, E, self.nu, alpha, False)

    # slp.jl:122-147
    def compute_derivative(self):
        return self.optimizer.merit_derivative(self.nu, self.feasibility_restoration)

    def norm_violations(self, p=1):                                         # slp.jl:174-178
        return self.optimizer.norm_violations(None, self.x, p) if self.optimizer is not None else \
            _host_violations(self.problem, self.E, self.x, p)

    def kt_residuals(self):                                                 # slp.jl:154
        if self.optimizer is None:                                          # before the first LP: lambda = 0
            return float(np.linalg.norm(self.df - self.mult_x_U - self.mult_x_L) /
                         max(1.0, np.linalg.norm(self.df)))
        return self.optimizer.kt_residuals(self.lam, self.mult_x_U, self.mult_x_L)

    def norm_complementarity(self):                                         # slp.jl:161-166
        if self.optimizer is None:
            return 0.0
        return self.optimizer.norm_complementarity(self.lam)

    def finish(self):                                                       # slp_line_search.jl:207-214
        pr = self.problem
        pr.obj_val = pr.eval_f(self.x)
        pr.status = int(self.ret)
        pr.x[:] = self.x
        if self.options.device_evaluator:      # the loop kept g on the device: one host evaluation for the write-back
            pr.eval_g(self.x, self.E)
        pr.g[:] = self.E
        pr.mult_g[:] = self.lam
        pr.mult_x_U[:] = self.mult_x_U
        pr.mult_x_L[:] = self.mult_x_L
        self.obj_val = pr.obj_val
        self.status = pr.status

    def _print(self, extra=""):
        if self.options.OutputFlag:
            print(f"{self.iter:6d}  {self.f: .8e}  {self.phi: .8e}  {self.directional_derivative: .4e}  "
                  f"{float(np.max(np.abs(self.p))):.4e}  {self.alpha:.4e}  {self.prim_infeas:.4e}  "
                  f"{self.dual_infeas:.4e}  {self.compl:.4e}  {self.lp_time:7.2f}{extra}")


def _host_violations(pr, E, x, p):
    viol = np.concatenate([
        np.where(E > pr.g_U, E - pr.g_U, np.where(E < pr.g_L, pr.g_L - E, 0.0)),
        np.where(x > pr.x_U, x - pr.x_U, np.where(x < pr.x_L, pr.x_L - x, 0.0))])
    if len(viol) == 0:
        return 0.0
    return float(np.max(np.abs(viol))) if p == INF else float(np.linalg.norm(viol, p))


class SlpLS(_Slp):
    """Line-search SLP (slp_line_search.jl)."""

    def compute_nu(self):                                                   # :251-261
        if self.iter == 1:
            self.nu = np.abs(self.lam)
        else:
            self.nu = np.maximum(self.nu, np.abs(self.lam))

    def compute_alpha(self):                                                # :222-244
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

    def run(self):                                                          # :78-215
        o = self.options
        self.start_time = time.time()
        self.clip_start()
        self.iter = 1
        if self.optimizer is None:
            self.optimizer = self._instantiate()
        dev_eval = bool(o.device_evaluator)
        if dev_eval:
            self.optimizer.attach_acopf(self.problem.source)
        while True:
            self.alpha = 0.0
            # KKT metrics with the multipliers of the previous iteration (App. C-7); they need the
            # Jacobian of *this* iterate on the device, which the update below provides
            if dev_eval:
                self.optimizer.eval_acopf(self.x, 1000.0, self.feasibility_restoration)
                self.f = float(np.atleast_1d(self.optimizer.get_eval_f())[0])
            else:
                self.eval_functions()
                self.optimizer.update(self.x, self.f, self.df, self.E, self.dE, 1000.0, self.feasibility_restoration)
            self.prim_infeas = self.optimizer.norm_violations(None, None, INF)
            self.dual_infeas = self.kt_residuals()
            self.compl = self.norm_complementarity()
            self.p, self.lam, self.mult_x_U, self.mult_x_L, self.p_slack, status = self.sub_optimize(1000.0, True)
            if status not in (LP_OPTIMAL, LP_INFEASIBLE):
                # Deviation from slp_line_search.jl:129, whose `slp.ret == -3` is a comparison, not an assignment: the
                # run would end as -5 "Optimize_not_called" after a full solve.  -3 (Error_In_Step_Computation) is the
                # evident intent.
                self.ret = -3
                if self.prim_infeas <= o.tol_infeas:
                    self.ret = 6
                break
            elif status == LP_INFEASIBLE:
                if self.feasibility_restoration:
                    self.ret = 6 if self.prim_infeas <= o.tol_infeas else 2
                    break
                self.feasibility_restoration = True
                continue
            self.compute_nu()
            self.phi = self.compute_phi(self.x, 0.0, self.p)
            self.directional_derivative = self.compute_derivative()
            is_valid_step = self.compute_alpha()
            self._print()
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
    """Trust-region SLP (slp_trust_region.jl)."""

    def __init__(self, model):
        super().__init__(model)
        self.delta = self.options.tr_size                                   # :62-65
        self.delta_max = 2.0
        self.alpha1 = 0.1
        self.alpha2 = 0.25

    def compute_nu(self):                                                   # slp.jl:54-66
        if self.iter == 1:
            norm_df = 1.0 if self.feasibility_restoration else float(np.linalg.norm(self.df))
            rn = self.optimizer.row_norms()
            self.nu = np.maximum(1.0, norm_df / np.maximum(1.0, rn))
        else:
            self.nu = np.maximum(self.nu, np.abs(self.lam))

    def step_quality(self):                                                 # :213-251
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

    def run(self):                                                          # :87-206
        o = self.options
        pr = self.problem
        self.start_time = time.time()
        self.clip_start()
        self.iter = 1
        while True:
            self.eval_functions()
            self.p, self.lam, self.mult_x_U, self.mult_x_L, self.p_slack, status = self.sub_optimize(self.delta)
            if status not in (LP_OPTIMAL, LP_INFEASIBLE):
                self.ret = -3                      # same deviation as in SlpLS.run (slp_trust_region.jl:131)
                if self.optimizer.norm_violations(pr.eval_g(self.x, np.zeros(pr.m)), self.x, 1) <= o.tol_infeas:
                    self.ret = 6
                break
            elif status == LP_INFEASIBLE:
                if self.feasibility_restoration:
                    self.ret = 6 if self.prim_infeas <= o.tol_infeas else 2
                    break
                self.feasibility_restoration = True
                continue
            self.compute_nu()
            self.prim_infeas = self.optimizer.norm_violations(None, None, INF)
            self.dual_infeas = self.kt_residuals()
            self.compl = self.norm_complementarity()
            if self.prim_infeas <= o.tol_infeas and self.compl <= o.tol_residual and \
                    float(np.max(np.abs(self.p))) <= o.tol_direction:
                if self.feasibility_restoration:
                    self.feasibility_restoration = False
                    self.iter += 1
                    # Deviation from slp_trust_region.jl:163-170, which `continue`s here without ever reaching
                    # its max_iter test: once the trust region has collapsed (delta ~ 1e-7) at a point that is
                    # feasible to tolerance but whose linearisation is infeasible inside the tiny box, the
                    # reference alternates normal LP (INFEASIBLE) / restoration LP forever.  Stop as its own
                    # max_iter branch (:177-183) would.
                    if self.iter >= o.max_iter:
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
            self._print(f"  rho {rho: .3e} delta {self.delta:.3e}")
            if self.ret in (0, 2, 6):
                break
            if rho >= 0:
                self.x = self.x + self.p
            self.iter += 1
        self.finish()
        return self


class SlpLSBatch:
    """B independent line-search SLP solves that share one Jacobian pattern (load scenarios of one network,
    BASELINE config 5), advanced in lock-step so that every round is **one** batched call of the device hot path:
    one `update`, the KKT reductions, one sub-LP batch solve (two when some scenarios are in feasibility
    restoration), and one batched merit evaluation per backtracking round.  Per scenario the control flow is
    exactly ``SlpLS.run`` (reference ``slp_line_search.jl:78-215``) — a scenario that terminates simply stops
    moving while the others go on.  ``problems`` are objects with the ``Model.from_problem`` interface.

    ``device_evaluator=True`` (ACOPF scenarios of one network, B200 engine): the NLP callbacks are not called at all —
    f, g and the Jacobian are evaluated by the device-side ACOPF evaluator (``SubLp.eval_acopf``) and every
    backtracking trial is one ``SubLp.acopf_trial`` call for the whole batch."""

    def __init__(self, problems, parameters: Parameters | None = None, device_evaluator: bool = False):
        self.device_evaluator = bool(device_evaluator)
        self.problems = list(problems)
        self.options = parameters if parameters is not None else Parameters()
        p0 = self.problems[0]
        self.B, self.n, self.m = len(self.problems), p0.n, p0.m
        for pr in self.problems:
            assert pr.n == self.n and pr.m == self.m and np.array_equal(pr.j_str, p0.j_str), \
                "batched scenarios must share the Jacobian pattern"
        B, n, m = self.B, self.n, self.m
        self.x = np.array([np.array(pr.x0, dtype=float) for pr in self.problems])
        self.f = np.zeros(B)
        self.df = np.zeros((B, n))
        self.E = np.zeros((B, m))
        self.dE = np.zeros((B, len(p0.j_str)))
        self.p = np.zeros((B, n))
        self.lam = np.zeros((B, m))
        self.mult_x_U = np.zeros((B, n))
        self.mult_x_L = np.zeros((B, n))
        self.nu = np.zeros((B, m))
        self.iter = np.ones(B, dtype=np.int64)
        self.ret = np.full(B, -5, dtype=np.int64)
        self.running = np.ones(B, dtype=bool)
        self.fr = np.zeros(B, dtype=bool)
        self.alpha = np.zeros(B)
        self.prim_infeas = np.full(B, INF)
        self.dual_infeas = np.full(B, INF)
        self.compl = np.full(B, INF)
        self.obj_val = np.zeros(B)
        self.rounds = 0
        self.lp_iterations = 0
        self.optimizer = None

    def _instantiate(self):
        o = self.options
        prs = self.problems
        arr = lambda name: np.array([np.asarray(getattr(pr, name), dtype=float) for pr in prs])  # noqa: E731
        if o.external_optimizer == "B200LP":
            opt = SubLp(self.n, self.m, prs[0].j_str, arr("x_L"), arr("x_U"), arr("g_L"), arr("g_U"), batch=self.B,
                        device=o.device, **{"warm_start": 1, **o.lp_options})
            opt._squeeze = False
            if self.device_evaluator:
                # the device evaluator holds ONE network: every scenario must share everything but its loads (which
                # only enter g_L / g_U)
                n0 = prs[0].net
                for k, pr in enumerate(prs[1:], 1):
                    for name in ("f_bus", "t_bus", "br_r", "br_x", "br_b", "tap", "shift", "gs", "bs", "cost2", "cost1",
                                 "cost0", "gen_bus", "dc_loss1"):
                        if not np.array_equal(getattr(pr.net, name), getattr(n0, name)):
                            raise ValueError(f"device_evaluator: scenario {k} differs from scenario 0 in network.{name}")
                    if pr.net.ref_bus != n0.ref_bus:
                        raise ValueError(f"device_evaluator: scenario {k} has another reference bus")
                opt.attach_acopf(prs[0])
            return opt
        if self.device_evaluator:
            raise ValueError("device_evaluator needs the B200 engine")
        return o.external_optimizer(self.n, self.m, prs[0].j_str, arr("x_L"), arr("x_U"), arr("g_L"), arr("g_U"),
                                    batch=self.B)

    def _eval(self, s):
        pr = self.problems[s]
        self.f[s] = pr.eval_f(self.x[s])
        pr.eval_grad_f(self.x[s], self.df[s])
        pr.eval_g(self.x[s], self.E[s])
        pr.eval_jac_g(self.x[s], "eval", None, None, self.dE[s])

    def _solve_phase(self, fr, sel=None):
        """Sub-LP batch of one phase; returns the extract tuple plus phi(0) and the directional derivative computed
        while the device still holds this phase's step and slacks.  ``sel``: the scenarios that are running in this
        phase -- the others (terminated, or in the other phase, whose normal LP is infeasible by construction) are
        masked out of the solve instead of being re-solved every round."""
        opt = self.optimizer
        if sel is not None and hasattr(opt, "set_active"):
            opt.set_active(sel)
        if fr and self.device_evaluator:
            opt.eval_acopf(self.x, 1000.0, True)
        elif fr:                    # the normal-phase data push was done for the KKT metrics of this round
            opt.update(self.x, self.f, self.df, self.E, self.dE, 1000.0, True)
        p, lam, mu_u, mu_l, slack, status = opt.solve_extract()
        info = getattr(opt, "last_info", None) or []
        self.lp_iterations += int(sum(i.get("iterations") or 0 for i in info))
        return [np.atleast_2d(a) if np.ndim(a) < 2 else a for a in (p, lam, mu_u, mu_l)] + [slack, np.atleast_1d(status)]

    def run(self):
        o = self.options
        B = self.B
        for s, pr in enumerate(self.problems):                              # slp_line_search.jl:96-105
            lo = pr.x_L > -INF
            self.x[s][lo] = np.maximum(self.x[s][lo], pr.x_L[lo])
            up = pr.x_U > -INF
            self.x[s][up] = np.minimum(self.x[s][up], pr.x_U[up])
        if self.optimizer is None:
            self.optimizer = self._instantiate()
        opt = self.optimizer
        while self.running.any():
            self.rounds += 1
            run = np.nonzero(self.running)[0]
            # KKT metrics with last round's multipliers on this round's Jacobian (App. C-7)
            if self.device_evaluator:
                opt.eval_acopf(self.x, 1000.0, False)
                self.f[:] = opt.get_eval_f()
            else:
                for s in run:
                    self._eval(s)
                opt.update(self.x, self.f, self.df, self.E, self.dE, 1000.0, False)
            self.prim_infeas[run] = np.atleast_1d(opt.norm_violations(None, None, INF))[run]
            self.dual_infeas[run] = np.atleast_1d(opt.kt_residuals(self.lam, self.mult_x_U, self.mult_x_L))[run]
            self.compl[run] = np.atleast_1d(opt.norm_complementarity(self.lam))[run]
            need = {False: self.running & ~self.fr, True: self.running & self.fr}
            status = np.zeros(B, dtype=np.int64)
            phi0 = np.zeros(B)
            deriv = np.zeros(B)
            for fr in (False, True):
                sel = need[fr]
                if not sel.any():
                    continue
                p, lam, mu_u, mu_l, slack, st = self._solve_phase(fr, sel)
                idx = np.nonzero(sel)[0]
                self.p[idx], self.lam[idx], self.mult_x_U[idx], self.mult_x_L[idx] = p[idx], lam[idx], mu_u[idx], mu_l[idx]
                status[idx] = st[idx]
                ok = idx[st[idx] == LP_OPTIMAL]
                # compute_nu! (:251-261) then phi(0) and D with this phase's device state
                first = ok[self.iter[ok] == 1]
                later = ok[self.iter[ok] != 1]
                self.nu[first] = np.abs(self.lam[first])
                self.nu[later] = np.maximum(self.nu[later], np.abs(self.lam[later]))
                base = self.prim_infeas if fr else self.f
                phi0[sel] = np.atleast_1d(opt.merit_phi(base, None, self.nu, np.zeros(B), fr))[sel]
                deriv[sel] = np.atleast_1d(opt.merit_derivative(self.nu, fr))[sel]
                self._line_search(fr, ok, phi0, deriv)
            # ---- per-scenario control flow of run! (:127-205)
            for s in run:
                st = status[s]
                if st not in (LP_OPTIMAL, LP_INFEASIBLE):
                    self.ret[s] = 6 if self.prim_infeas[s] <= o.tol_infeas else -3   # see SlpLS.run
                    self.running[s] = False
                    continue
                if st == LP_INFEASIBLE:
                    if self.fr[s]:
                        self.ret[s] = 6 if self.prim_infeas[s] <= o.tol_infeas else 2
                        self.running[s] = False
                    else:
                        self.fr[s] = True                                   # re-solved next round, iter unchanged
                    continue
                if self.iter[s] >= o.max_iter:
                    self.ret[s] = 6 if self.prim_infeas[s] <= o.tol_infeas else -1
                    self.running[s] = False
                    continue
                if (self.prim_infeas[s] <= o.tol_infeas and self.compl[s] <= o.tol_residual) or \
                        float(np.max(np.abs(self.p[s]))) <= o.tol_direction:
                    if self.fr[s]:
                        self.fr[s] = False
                        self.iter[s] += 1
                        continue
                    elif self.dual_infeas[s] <= o.tol_residual:
                        self.ret[s] = 0
                        self.running[s] = False
                        continue
                if not self._valid[s]:
                    if self._ret3[s]:
                        self.ret[s] = 6 if self.prim_infeas[s] <= o.tol_infeas else 2
                        self.running[s] = False
                    else:
                        self.fr[s] = True
                        self.iter[s] += 1
                    continue
                self.x[s] = self.x[s] + self.alpha[s] * self.p[s]
                self.iter[s] += 1
        for s, pr in enumerate(self.problems):
            self.obj_val[s] = pr.eval_f(self.x[s])
        return self

    def _line_search(self, fr, idx, phi0, deriv):
        """compute_alpha (:222-244) for the scenarios ``idx`` of one phase, one batched merit call per trial round."""
        o = self.options
        B = self.B
        if not hasattr(self, "_valid"):
            self._valid = np.ones(B, dtype=bool)
            self._ret3 = np.zeros(B, dtype=bool)
        self._valid[idx] = True
        self._ret3[idx] = False
        self.alpha[idx] = 1.0
        searching = np.zeros(B, dtype=bool)
        searching[idx] = True
        base = np.zeros(B)
        Et = self.E.copy()
        while searching.any():
            act = np.nonzero(searching)[0]
            if self.device_evaluator:
                phi = np.atleast_1d(self.optimizer.acopf_trial(self.alpha, self.nu, self.prim_infeas if fr else None, fr))
            else:
                for s in act:
                    pr = self.problems[s]
                    xt = self.x[s] + self.alpha[s] * self.p[s]
                    pr.eval_g(xt, Et[s])
                    base[s] = self.prim_infeas[s] if fr else pr.eval_f(xt)
                phi = np.atleast_1d(self.optimizer.merit_phi(base, Et, self.nu, self.alpha, fr))
            for s in act:
                if phi[s] > phi0[s] + o.eta * self.alpha[s] * deriv[s]:
                    if self.alpha[s] < o.min_alpha:
                        if fr:
                            self._ret3[s] = True
                        self._valid[s] = False
                        searching[s] = False
                    else:
                        self.alpha[s] *= o.tau
                else:
                    searching[s] = False
