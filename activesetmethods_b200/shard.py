"""Scenario sharding across GPUs: independent load scenarios share one sparsity pattern and are
split into contiguous blocks, one process per GPU, with **no collective on the data path**.  ``torch.distributed``
is used only to time (max over ranks) and to gather the per-scenario results (status, objective, iterations) —
the final gather the reference would do over its own loop of JuMP models (examples/acopf/opf.jl:23-36 per scenario).
"""
from __future__ import annotations

import numpy as np


def scenario_ids(rank: int, world: int, per_gpu: int, first: int = 1):
    """Contiguous block of scenario ids (= RNG seeds) owned by ``rank``; weak scaling."""
    if not (0 <= rank < world) or per_gpu < 0:
        raise ValueError("bad rank / world / per_gpu")
    return [first + rank * per_gpu + s for s in range(per_gpu)]


def split_scenarios(total: int, world: int, first: int = 1):
    """Strong-scaling split of ``total`` scenarios: contiguous blocks whose sizes differ by at most one."""
    base, extra = divmod(total, world)
    out, start = [], first
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append(list(range(start, start + cnt)))
        start += cnt
    return out


def row_blocks(row_ptr, world: int):
    """Contiguous row blocks balanced by nonzeros (+2 per row) for the row-partitioned single instance
    (SURVEY.md §8e): returns ``world + 1`` row offsets."""
    rp = np.asarray(row_ptr, dtype=np.int64)
    m = len(rp) - 1
    weight = rp + 2 * np.arange(m + 1)
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(weight, weight[-1] * r / world)))
    cuts.append(m)
    return [min(max(c, cuts[i - 1] if i else 0), m) for i, c in enumerate(cuts)]


def max_over_ranks(value: float, device=None) -> float:
    """Device-side time of a step = the slowest rank's."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_results(ids, status, objective, iterations, device=None):
    """All ranks contribute equally sized blocks; every rank gets ``(ids, status, objective, iterations)`` of the
    whole job ordered by scenario id."""
    import torch
    import torch.distributed as dist
    mine = torch.tensor(np.stack([np.asarray(ids, dtype=np.float64), np.asarray(status, dtype=np.float64),
                                  np.asarray(objective, dtype=np.float64), np.asarray(iterations, dtype=np.float64)]),
                        dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        allr = mine.cpu().numpy()
    else:
        parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, mine)
        allr = torch.cat(parts, dim=1).cpu().numpy()
    order = np.argsort(allr[0], kind="stable")
    allr = allr[:, order]
    return allr[0].astype(np.int64), allr[1].astype(np.int32), allr[2], allr[3].astype(np.int64)
