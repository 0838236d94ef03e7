"""Development prototype (numpy) of the restarted reflected-Halpern PDHG LP solver that
``activesetmethods_b200/csrc`` implements in CUDA.  Not part of the product and not the oracle: it exists so
algorithmic choices (scaling, restart rule, primal-weight control, infeasibility test) can be tuned on the CPU
against HiGHS before they are frozen into kernels.

LP form:  min c'x  s.t.  rl <= Kx <= ru,  lb <= x <= ub   (entries may be +-inf).
"""
from __future__ import annotations

import math
import numpy as np
import scipy.sparse as sp

INF = math.inf


def ruiz_pc_scale(K: sp.csr_matrix, ruiz_iters=10, pc=True, tiny_rel=1e-8):
    m, n = K.shape
    dr = np.ones(m)
    dc = np.ones(n)
    # rows / columns whose largest coefficient is below tiny_rel * max|K| are numerically empty: they keep scale 1
    # (kTinyRel in csrc/lp_solver.cuh); they are masked out while the factors are computed and restored at the end
    a0 = abs(K).tocsr()
    gmax = a0.max() if K.nnz else 0.0
    keep_r = np.asarray(a0.max(axis=1).todense()).ravel() > tiny_rel * gmax
    keep_c = np.asarray(a0.max(axis=0).todense()).ravel() > tiny_rel * gmax
    A = (sp.diags(keep_r.astype(float)) @ K @ sp.diags(keep_c.astype(float))).tocsr()
    for _ in range(ruiz_iters):
        absA = abs(A)
        rmax = np.asarray(absA.max(axis=1).todense()).ravel()
        cmax = np.asarray(absA.max(axis=0).todense()).ravel()
        sr = np.where(rmax > 0, 1.0 / np.sqrt(rmax), 1.0)
        sc = np.where(cmax > 0, 1.0 / np.sqrt(cmax), 1.0)
        A = sp.diags(sr) @ A @ sp.diags(sc)
        dr *= sr
        dc *= sc
    if pc:
        absA = abs(A)
        r1 = np.asarray(absA.sum(axis=1)).ravel()
        c1 = np.asarray(absA.sum(axis=0)).ravel()
        sr = np.where(r1 > 0, 1.0 / np.sqrt(r1), 1.0)
        sc = np.where(c1 > 0, 1.0 / np.sqrt(c1), 1.0)
        A = sp.diags(sr) @ A @ sp.diags(sc)
        dr *= sr
        dc *= sc
    A = sp.diags(dr) @ K @ sp.diags(dc)
    return A.tocsr(), dr, dc


def power_norm(A, iters=40, seed=0):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(A.shape[1])
    v /= np.linalg.norm(v)
    s = 1.0
    for _ in range(iters):
        w = A @ v
        v = A.T @ w
        s = np.linalg.norm(v)
        if s == 0:
            return 0.0
        v /= s
    return math.sqrt(s)


def solve(K, c, lb, ub, rl, ru, eps=1e-6, max_iter=200000, check_every=64, verbose=False,
          x0=None, y0=None, kp=0.99, ki=0.96, kd=0.0, reflect=1.0, eps_infeas=1e-8, omega0=None,
          bound_obj_rescale=True, ruiz_iters=10, pc=True, b_suf=0.2, b_nec=0.8, b_art=0.36, clamp=None, i_smooth=1.0,
          guard=0.0, power=False):
    K = sp.csr_matrix(K)
    m, n = K.shape
    A, dr, dc = ruiz_pc_scale(K, ruiz_iters, pc)
    AT = A.T.tocsr()
    # scaled problem: x = dc * xs ; ys = y / dr  (y = dr * ys)
    cs = c * dc
    lbs, ubs = lb / dc, ub / dc
    rls, rus = rl * dr, ru * dr
    # bound-objective rescaling: put ||c|| and ||b|| at O(1)
    bfin = np.concatenate([rls[np.isfinite(rls)], rus[np.isfinite(rus) & (rus != rls)]])
    sb = 1.0 / (np.linalg.norm(bfin) + 1.0) if bound_obj_rescale else 1.0
    sc_ = 1.0 / (np.linalg.norm(cs) + 1.0) if bound_obj_rescale else 1.0
    cs = cs * sc_
    lbs, ubs, rls, rus = lbs * sb, ubs * sb, rls * sb, rus * sb
    # x_s = sb * x/dc ; y_s = sc_ * y/dr

    normA = power_norm(A, 400) * 1.01 if (not pc or power) else 1.0  # Pock-Chambolle (alpha=1) guarantees ||A||_2 <= 1; power iteration under-estimates
    eta = 0.998 / normA if normA > 0 else 1.0
    if verbose: print('normA', normA)
    nq = np.linalg.norm(np.concatenate([rls[np.isfinite(rls)], rus[np.isfinite(rus) & (rus != rls)]]))
    nc = np.linalg.norm(cs)
    omega = omega0 if omega0 else (nc / nq if (nc > 0 and nq > 0) else 1.0)

    # unscaled norms for termination
    q_un = np.concatenate([rl[np.isfinite(rl)], ru[np.isfinite(ru) & (ru != rl)]])
    nq_un = np.linalg.norm(q_un)
    nc_un = np.linalg.norm(c)

    x = np.zeros(n) if x0 is None else x0 / dc * sb
    x = np.clip(x, lbs, ubs)
    y = np.zeros(m) if y0 is None else y0 / dr * sc_
    xa, ya = x.copy(), y.copy()           # anchor
    k = 0                                 # iterations within the epoch
    total = 0
    r0 = None
    r_prev = INF
    e_sum = 0.0
    e_prev = 0.0
    n_restart = 0
    hist = []
    omega_init = omega; best_omega = omega; best_bal = INF
    status = "ITERATION_LIMIT"
    fin_lb, fin_ub = np.isfinite(lb), np.isfinite(ub)
    while total < max_iter:
        tau, sigma = eta / omega, eta * omega
        # PDHG operator at z = (x, y)
        g = cs - AT @ y
        xp = np.clip(x - tau * g, lbs, ubs)
        xbar = 2 * xp - x
        Axbar = A @ xbar
        t = Axbar - y / sigma
        yp = np.where(t < rls, sigma * (rls - t), np.where(t > rus, sigma * (rus - t), 0.0))
        total += 1
        k += 1
        do_check = (total % check_every == 0) or k == 1
        if do_check:
            dx, dy = xp - x, yp - y
            # fixed point error in the M-norm
            r2 = dx @ dx / tau - 2 * (dy @ (A @ dx)) + dy @ dy / sigma
            r = math.sqrt(max(r2, 0.0))
            # --- termination on (xp, yp) in the original space
            xo = np.clip(xp * dc / sb, lb, ub)
            yo = yp * dr / sc_
            Kx = K @ xo
            pres = np.linalg.norm(Kx - np.clip(Kx, rl, ru))
            rc = c - K.T @ yo
            # reduced costs only where the iterate sits on the bound (OR-tools PDLP
            # handle_some_primal_gradients_on_finite_bounds_as_residuals); the rest is dual residual
            rpos = np.where(fin_lb & (xp <= lbs), np.maximum(rc, 0), 0.0)
            rneg = np.where(fin_ub & (xp >= ubs), np.minimum(rc, 0), 0.0)
            dres = np.linalg.norm(rc - rpos - rneg)
            pobj = c @ xo
            ypos, yneg = np.maximum(yo, 0), np.minimum(yo, 0)
            with np.errstate(invalid="ignore"):
                dobj = (np.sum(np.where(ypos > 0, rl * ypos, 0.0)) + np.sum(np.where(yneg < 0, ru * yneg, 0.0))
                        + np.sum(np.where(rpos > 0, lb * rpos, 0.0)) + np.sum(np.where(rneg < 0, ub * rneg, 0.0)))
            gap = abs(pobj - dobj)
            hist.append((total, pres, dres, gap, r, omega))
            if verbose and (total % (check_every * 16) == 0):
                print(f"it {total:7d} pres {pres:.2e} dres {dres:.2e} gap {gap:.2e} pobj {pobj:.8e} r {r:.2e} w {omega:.2e} restarts {n_restart}")
            if pres <= eps * (1 + nq_un) and dres <= eps * (1 + nc_un) and gap <= eps * (1 + abs(pobj) + abs(dobj)):
                status = "OPTIMAL"
                x, y = xp, yp
                break
            # --- infeasibility: dy as a Farkas ray
            ray = dy * dr / sc_
            nr = np.linalg.norm(ray, np.inf)
            if nr > 0:
                ray = ray / nr
                rp, rn = np.maximum(ray, 0), np.minimum(ray, 0)
                bad = (np.any((rp > 1e-12) & ~np.isfinite(rl)) or np.any((rn < -1e-12) & ~np.isfinite(ru)))
                if not bad:
                    kty = K.T @ ray
                    with np.errstate(invalid="ignore"):
                        robj = (np.sum(np.where(rp > 0, rl * rp, 0.0)) + np.sum(np.where(rn < 0, ru * rn, 0.0)))
                        # box support: sum over j of min over x in [lb,ub] of -(kty_j) x_j
                        t = -kty
                        sup = np.where(t > 0, t * lb, np.where(t < 0, t * ub, 0.0))
                    if np.all(np.isfinite(sup)):
                        robj += np.sum(sup)
                        if robj > eps_infeas * max(1.0, np.linalg.norm(kty, np.inf)):
                            status = "INFEASIBLE"
                            x, y = xp, yp
                            break
        # Halpern + reflection
        w = (k) / (k + 1.0)
        xn = w * ((1 + reflect) * xp - reflect * x) + (1 - w) * xa
        yn = w * ((1 + reflect) * yp - reflect * y) + (1 - w) * ya
        if do_check:
            restart = False
            if k == 1:
                r0 = r
            else:
                if r <= b_suf * r0:
                    restart = True
                elif r <= b_nec * r0 and r > r_prev:
                    restart = True
                elif k >= b_art * total:
                    restart = True
            r_prev = r
            if restart:
                ddx = np.linalg.norm(xp - xa)
                ddy = np.linalg.norm(yp - ya)
                if guard > 0 and (ddx <= guard or ddy <= guard or omega < omega_init * 1e-5 or omega > omega_init * 1e5):
                    omega = best_omega
                    e_sum = 0.0
                    e_prev = 0.0
                elif ddx > 1e-300 and ddy > 1e-300:
                    e = math.log((math.sqrt(omega) * ddx) / (ddy / math.sqrt(omega)))
                    e_sum = i_smooth * e_sum + e
                    dlog = -(kp * e + ki * e_sum + kd * (e - e_prev))
                    if clamp:
                        dlog = max(-clamp, min(clamp, dlog))
                    omega = math.exp(math.log(omega) + dlog)
                    e_prev = e
                rp_ = pres / (1 + nq_un); rd_ = dres / (1 + nc_un)
                if rp_ > 0 and rd_ > 0:
                    bal = abs(math.log10(rd_ / rp_))
                    if bal < best_bal:
                        best_bal = bal; best_omega = omega
                xn, yn = xp.copy(), yp.copy()
                xa, ya = xn.copy(), yn.copy()
                k = 0
                r0 = None
                r_prev = INF
                n_restart += 1
        x, y = xn, yn
    xo = x * dc / sb
    yo = y * dr / sc_
    return dict(status=status, x=xo, y=yo, iters=total, restarts=n_restart, hist=hist, obj=float(c @ xo))
