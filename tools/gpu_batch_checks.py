"""Developer tool: robustness sweep of the barrier engine over scenario batches of several cases (statuses, Newton
steps, residuals, time) and full lock-step SLP solves of a scenario batch.
python tools/gpu_batch_checks.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from activesetmethods_b200.sublp import SubLp
from activesetmethods_b200.examples import acopf
from activesetmethods_b200.slp import Parameters, SlpLSBatch

for case, S, delta in (("case118", 1024, 1000.0), ("case2869pegase", 128, 1000.0), ("case1354pegase", 256, 0.4),
                       ("case13659pegase", 32, 1000.0)):
    net = bench.network(case)
    mdl, d = bench.linearise(net, list(range(1, S + 1)))
    lp = SubLp(mdl.n, mdl.m, mdl.j_str, d["xL"], d["xU"], d["gL"], d["gU"], batch=S)
    for fr in (False, True):
        for rep in range(2):
            t0 = time.time()
            out = lp.sub_optimize(d["x"], d["f"], d["df"], d["E"], d["dE"], delta, fr)
            dt = time.time() - t0
        st = np.array([i["status"] for i in lp.last_info])
        it = np.array([i["iterations"] for i in lp.last_info])
        ms, _ = lp.last_solve_timing()
        print(f"{case} B={S} delta={delta} fr={fr}: statuses {dict(zip(*np.unique(st, return_counts=True)))} newton {it.min()}..{it.max()} "
              f"device {ms:.0f} ms call {dt*1e3:.0f} ms -> {S/dt:.0f} LPs/s", flush=True)
    lp.close()

net = acopf.synthetic_network(*acopf.PEGASE_SHAPES["case118"])
for dev in (False, True):
    probs = [acopf.AcopfModel(acopf.perturb_loads(net, s)) for s in range(1, 65)]
    t0 = time.time()
    b = SlpLSBatch(probs, Parameters(max_iter=100), device_evaluator=dev).run()
    dt = time.time() - t0
    print(f"SlpLSBatch 64 x case118 (device evaluator {dev}): rounds {b.rounds} statuses {dict(zip(*np.unique(b.ret, return_counts=True)))} "
          f"SLP iterations {b.iter.min()}..{b.iter.max()} newton steps {b.lp_iterations} wall {dt:.1f} s -> {64/dt:.2f} solved scenarios/s", flush=True)
