"""Device time of launch-bound configurations (everything runs in the interpreter kernel):
cfg 1 (1000-step 3x3 chain), a balanced binary tree of 3x3 products, the 6x6 D=2 grid.
  python benchmarks/interp_latency.py"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contractn_b200 import engine, planner
from benchmarks.configs import CONFIGS


def run(name, einstr, params, optimize):
    shapes = tuple(tuple(p.shape) for p in params)
    lst = planner.cached_contract_list(einstr, shapes, optimize=optimize)
    plan = engine.Plan(lst, shapes, params[0].dtype)
    dev = [torch.as_tensor(p).cuda() for p in params]
    for _ in range(3):
        plan.execute_device(dev)
    best = None
    for _ in range(7):
        _, _, ms = plan.execute_device(dev, timed=True)
        best = sum(ms) if best is None else min(best, sum(ms))
    print(json.dumps({"case": name, "steps": len(lst), "kinds": sorted(set(plan.launch_kinds)),
                      "launches": len(plan.launch_kinds), "ms": round(best, 4),
                      "us_per_step": round(best * 1e3 / len(lst), 3)}), flush=True)


tn, optimize = CONFIGS["cfg1"][1]()
run("cfg1 chain 1000x(3,3)", tn.einsum_str, tn.params, optimize)

# balanced tree: 1024 matrices, explicit pairwise-halving path (depth 10)
n = 1024
rng = np.random.default_rng(5)
mats = [rng.uniform(0.5, 1.5, (3, 3)) for _ in range(n)]
syms = [planner.get_symbol(i) for i in range(n + 1)]
einstr = ",".join(syms[i] + syms[i + 1] for i in range(n)) + "->" + syms[0] + syms[n]
path, m = [], n
while m > 1:
    # contracting (0,1) appends the product at the end: one sweep over the level halves it in order
    path += [(0, 1)] * (m // 2)
    m = m - m // 2
run("balanced tree 1024x(3,3)", einstr, mats, tuple(path))

from benchmarks.configs import peps_tn
tn, optimize = peps_tn(6, 2)
run("grid 6x6 D=2", tn.einsum_str, tn.params, optimize)
