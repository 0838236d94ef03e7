"""Developer tool: B load scenarios of a case solved to the end by the lock-step batched SLP driver.
python tools/gpu_slp_batch.py case118 64 [max_iter] [dev]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.build()
import bench
from activesetmethods_b200.examples import acopf
from activesetmethods_b200.slp import Parameters, SlpLSBatch
case, B = sys.argv[1], int(sys.argv[2])
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 40
dev = len(sys.argv) > 4 and sys.argv[4] == "dev"       # device-side ACOPF evaluator instead of the host callbacks
net = bench.network(case)
probs = [acopf.AcopfModel(acopf.perturb_loads(net, s + 1)) for s in range(B)]
t0 = time.time()
b = SlpLSBatch(probs, Parameters(max_iter=max_iter, lp_options=dict(eps_rel=1e-6, max_iter=4000000)),
               device_evaluator=dev).run()
dt = time.time() - t0
print(f"{case} x {B}: {b.rounds} lock-step rounds, statuses {dict(zip(*np.unique(b.ret, return_counts=True)))}, SLP iterations "
      f"min/mean/max {b.iter.min()}/{b.iter.mean():.1f}/{b.iter.max()}, PDHG iterations {b.lp_iterations:.3g}, wall {dt:.1f}s "
      f"-> {B / dt:.2f} scenarios/s (full SLP solves), objective mean {b.obj_val.mean():.4f}")
