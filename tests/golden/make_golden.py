"""Generates the committed fixtures of tests/golden/.  Run in the build container (it reads the reference's
own data file, which does not exist on the GPU box):

    python tests/golden/make_golden.py

* ``case3_network.json``  — the network of /root/reference/examples/acopf/case3.m (the reference's only
  ACOPF known-answer input, test/opf.jl:6-22) parsed by ``acopf.parse_matpower`` into per-unit arrays.
* ``sublp_<name>.npz``    — for toy / case3 / case9: the linearisations (x, f, df, E, dE, delta, fr) of the first
  sub-LPs of the oracle's SLP run with the oracle's simplex answers (status, objective), so that GPU parity
  tests have committed vectors besides the live oracle.
"""
import dataclasses
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from activesetmethods_b200.examples import acopf, small_nlps  # noqa: E402
from oracle import slp_oracle as so  # noqa: E402


def network_to_json(net):
    d = {}
    for f in dataclasses.fields(net):
        v = getattr(net, f.name)
        d[f.name] = v.tolist() if isinstance(v, np.ndarray) else v
    return d


def main():
    net = acopf.parse_matpower(open("/root/reference/examples/acopf/case3.m").read())
    with open(os.path.join(HERE, "case3_network.json"), "w") as fh:
        json.dump(network_to_json(net), fh, indent=1)
    problems = {
        "toy": (small_nlps.ToyNlp(), "Line Search", 12),
        "case3": (acopf.AcopfModel(net), "Line Search", 6),
        "case9": (acopf.AcopfModel(acopf.case9()), "Line Search", 8),
        "case9tr": (acopf.AcopfModel(acopf.case9()), "Trust Region", 8),
    }
    for name, (pr, alg, limit) in problems.items():
        cls = so.SlpLS if alg == "Line Search" else so.SlpTR
        slp = cls(pr, so.Parameters(algorithm=alg, max_iter=60))
        lps = []
        slp.record = lambda s, d: lps.append(d)
        slp.run()
        lps = lps[:limit]
        log = slp.lp_log[:limit]
        np.savez_compressed(
            os.path.join(HERE, f"sublp_{name}.npz"),
            x=np.array([d["x"] for d in lps]), f=np.array([d["f"] for d in lps]),
            df=np.array([d["df"] for d in lps]), E=np.array([d["E"] for d in lps]),
            dE=np.array([d["dE"] for d in lps]), delta=np.array([d["delta"] for d in lps]),
            fr=np.array([d["fr"] for d in lps]), status=np.array([l[0] for l in log]),
            objective=np.array([np.nan if l[1] is None else l[1] for l in log]),
            final_status=slp.ret, final_objective=slp.obj_val, final_x=slp.x, final_iter=slp.iter)
        print(name, "LPs", len(lps), "final", slp.ret, slp.obj_val, [l[0] for l in log])


if __name__ == "__main__":
    main()
