"""N > 1 host logic on the CPU (gloo, world_size 2): scenario blocks are disjoint and complete, the result gather
orders by scenario id, timing takes the max over ranks.  The per-scenario "solve" is the CPU oracle on case9 load
scenarios (tests may use the oracle); no GPU is involved."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from activesetmethods_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, per_gpu, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from activesetmethods_b200.examples import acopf
    from oracle import slp_oracle as so
    ids = shard.scenario_ids(rank, world, per_gpu)
    net = acopf.case9()
    st, obj, its = [], [], []
    for sid in ids:
        mdl = acopf.AcopfModel(acopf.perturb_loads(net, sid))
        x = np.clip(mdl.x0, mdl.x_L, mdl.x_U)
        pat = so.JacobianPattern(mdl.m, mdl.n, mdl.j_str)
        ref = so.SubLp(pat, mdl.g_L, mdl.g_U, mdl.x_L, mdl.x_U)
        out = ref.solve(pat.assemble(mdl.eval_jac_g(x, "eval", None, None, np.zeros(mdl.nnz))),
                        mdl.eval_grad_f(x, np.zeros(mdl.n)), mdl.eval_f(x), mdl.eval_g(x, np.zeros(mdl.m)), x, 1000.0)
        st.append(out[5]); obj.append(ref.last_objective); its.append(sid * 10)
    t_max = shard.max_over_ranks(1.0 + rank)
    total = shard.sum_over_ranks(float(len(ids)))
    g_ids, g_st, g_obj, g_its = shard.gather_results(ids, st, obj, its)
    if rank == 0:
        q.put((t_max, total, g_ids.tolist(), g_st.tolist(), g_obj.tolist(), g_its.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding():
    world, per_gpu = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, per_gpu, q)) for r in range(world)]
    for p in procs:
        p.start()
    t_max, total, ids, st, obj, its = q.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert t_max == 2.0 and total == world * per_gpu
    assert ids == list(range(1, 1 + world * per_gpu))
    assert all(s == 0 for s in st) and its == [10 * i for i in ids]
    assert len(set(np.round(obj, 6))) == len(obj)          # different load scenarios -> different optima


def test_block_helpers():
    assert shard.scenario_ids(1, 4, 128) == list(range(129, 257))
    parts = shard.split_scenarios(10, 4)
    assert [len(p) for p in parts] == [3, 3, 2, 2] and sum(parts, []) == list(range(1, 11))
    assert shard.max_over_ranks(3.5) == 3.5            # no process group: identity


def test_strong_scaling_split_is_the_bench_partition():
    """bench.py's job: 1024 scenarios split over 1 / 2 / 4 / 8 ranks -- contiguous, disjoint, complete, equal sizes."""
    for world in (1, 2, 4, 8):
        parts = shard.split_scenarios(1024, world)
        assert len(parts) == world and all(len(p) == 1024 // world for p in parts)
        assert sum(parts, []) == list(range(1, 1025))
    parts = shard.split_scenarios(1000, 8)                 # not divisible: sizes differ by at most one
    assert sorted(set(len(p) for p in parts)) == [125] and sum(parts, []) == list(range(1, 1001))
    parts = shard.split_scenarios(1001, 8)
    assert max(len(p) for p in parts) - min(len(p) for p in parts) == 1 and sum(len(p) for p in parts) == 1001
