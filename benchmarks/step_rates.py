"""Device time of single non-GEMM steps (HBM-bound classes): GB/s vs measured HBM peak."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contractn_b200 import engine, planner

CASES = [
    ("bps,bp->bs", [(4096, 2, 512), (4096, 2)], np.float32),
    ("bps,bp->bs", [(4096, 2, 512), (4096, 2)], np.float64),
    ("ca,ba->cab", [(1024, 256), (1024, 256)], np.float64),
    ("ab,ab->ab", [(2, 1 << 20), (2, 1 << 20)], np.float64),
    ("ab,bc->ac", [(4096, 2), (2, 512)], np.float32),
    ("ab,ab->", [(8192, 8192), (8192, 8192)], np.float64),
    ("abcdefg,gx->abcdefx", [(8,) * 7, (8, 8)], np.float64),
]
for einstr, shapes, dt in CASES:
    tdt = torch.float32 if dt == np.float32 else torch.float64
    ops = [torch.rand(s, dtype=tdt, device="cuda") for s in shapes]
    lst = planner.contract_path(einstr, *shapes)[1]
    plan = engine.Plan(lst, shapes, dt)
    for _ in range(3):
        plan.execute_device(ops)
    best = None
    for _ in range(5):
        _, _, ms = plan.execute_device(ops, timed=True)
        best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
    print(json.dumps({"einsum": einstr, "dtype": np.dtype(dt).name, "kinds": plan.launch_kinds,
                      "us": [round(x * 1e3, 1) for x in best],
                      "GBps_first": round(plan.step_bytes / best[0] / 1e6, 1)}), flush=True)
