"""Worker of tests/test_gpu_dist.py (run under torchrun, one rank per GPU): the first sub-LP of a case solved as a
row-partitioned instance over all ranks, checked against the oracle's simplex solve and the single-GPU engine."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from helpers import problem
    from oracle import slp_oracle as so
    from activesetmethods_b200.sublp import B200LP, B200RowPartitionedLP, nccl_unique_id

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    name = sys.argv[1] if len(sys.argv) > 1 else "case118"
    eps = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-7
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")          # host-side plumbing only: the data path uses the library's own NCCL comm
    pr = problem(name)
    x = np.clip(pr.x0, pr.x_L, pr.x_U)
    pat = so.JacobianPattern(pr.m, pr.n, pr.j_str)
    ref = so.SubLp(pat, pr.g_L, pr.g_U, pr.x_L, pr.x_U)
    vals = pat.assemble(pr.eval_jac_g(x, "eval", None, None, np.zeros(len(pr.j_str))))
    K, cost, off, lb, ub, rl, ru = ref.build(vals, pr.eval_grad_f(x, np.zeros(pr.n)), pr.eval_f(x),
                                             pr.eval_g(x, np.zeros(pr.m)), x, 1000.0, False)
    K = K[:pr.m, :pr.n].tocsr()
    K.sort_indices()
    cost, lb, ub, rl, ru = cost[:pr.n], lb[:pr.n], ub[:pr.n], rl[:pr.m], ru[:pr.m]
    ids = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    lp = B200RowPartitionedLP(pr.n, pr.m, K.indptr, K.indices, rank, world, ids[0], device=local, eps_rel=eps,
                              max_iter=4_000_000)
    lp.set_matrix_values(K.data)
    lp.set_objective(cost, off)
    lp.set_col_bounds(lb, ub)
    lp.set_row_bounds(rl, ru)
    info = lp.optimize()[0]
    xs = lp.primal()
    y_loc = lp.row_dual()
    parts = [None] * world
    dist.all_gather_object(parts, (lp.r0, lp.r1, y_loc))
    ok = True
    if rank == 0:
        y = np.concatenate([p[2] for p in sorted(parts, key=lambda t: t[0])])
        st, xr, yr, dr, obj = so.HighsLp().solve(K.tocsc(), cost, off, lb, ub, rl, ru)
        one = B200LP(pr.n, pr.m, K.indptr, K.indices, device=local, eps_rel=eps, max_iter=4_000_000)
        one.set_matrix_values(K.data); one.set_objective(cost, off); one.set_col_bounds(lb, ub); one.set_row_bounds(rl, ru)
        i1 = one.optimize()[0]
        Kx = K @ xs
        feas = max(np.max(np.maximum(0, rl - Kx)), np.max(np.maximum(0, Kx - ru)), np.max(np.maximum(0, lb - xs)),
                   np.max(np.maximum(0, xs - ub)))
        rel = abs(info["objective"] - obj) / max(1.0, abs(obj))
        rel1 = abs(i1["objective"] - obj) / max(1.0, abs(obj))
        rc = cost - K.T @ y
        print(f"DIST {name} world {world}: status {info['status']} objective {info['objective']:.9f} vs simplex {obj:.9f} "
              f"(rel {rel:.2e}; single GPU rel {rel1:.2e}), primal infeasibility {feas:.2e}, iterations "
              f"{info['iterations']} (single GPU {i1['iterations']}), |y| {np.linalg.norm(y):.6e}", flush=True)
        ok = info["status"] == 0 and st == 0 and rel <= 1e-6 and feas <= 1e-6 and len(y) == pr.m and np.all(np.isfinite(rc))
        print("DIST_OK" if ok else "DIST_FAIL", flush=True)
    dist.barrier()
    lp.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
