"""Per-launch device times of one configuration, heaviest first.
  python benchmarks/profile_plan.py cfg5 [top_n]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contractn_b200 import engine, planner
from benchmarks.configs import CONFIGS

name = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 15
_, build, _ = CONFIGS[name]
tn, optimize = build()
einstr, params = tn.einsum_str, tn.params
shapes = tuple(tuple(p.shape) for p in params)
lst = planner.cached_contract_list(einstr, shapes, optimize=optimize)
plan = engine.Plan(lst, shapes, params[0].dtype, flags=engine.CTN_PLAN_NO_INTERPRETER)
dev = [torch.as_tensor(p).cuda() for p in params]
for _ in range(2):
    plan.execute_device(dev)
_, _, ms = plan.execute_device(dev, timed=True)
# with the interpreter off and no chain fusion in these configs, compute launches map 1:1 to steps
steps = [c[2] for c in lst]
rows, si = [], 0
for kind, t in zip(plan.launch_kinds, ms):
    if kind in ("finalize", "materialize", "gemm_reduce"):
        rows.append((t, kind, "-"))
    else:
        rows.append((t, kind, steps[si] if si < len(steps) else "?"))
        si += 1
total = sum(ms)
print(f"{name}: {len(ms)} launches, {total:.3f} ms")
sizes = {}
for term, shp in zip(einstr.split("->")[0].split(","), shapes):
    for ch, dd in zip(term, shp):
        sizes[ch] = dd
def vol(s):
    v = 1
    for ch in s: v *= sizes[ch]
    return v
for t, kind, st in sorted(rows, reverse=True)[:top]:
    extra = ""
    if st not in ("-", "?"):
        ins, out = st.split("->"); l, r = ins.split(",")
        k = [ch for ch in set(l) & set(r) if ch not in out]
        extra = f" lhs={vol(l):.3g} rhs={vol(r):.3g} out={vol(out):.3g} K={vol(k):.3g} len=({len(l)},{len(r)})->{len(out)}"
    print(f"{t:9.3f} ms {100*t/total:5.1f}%  {kind:8s}{extra}")
