"""Bring-up check of the tcgen05 3xTF32 GEMM kernel: progressively harder cases,
each printed as soon as it finishes (run under `timeout`)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contractn_b200 import engine, planner
import torch

CASES = [
    ("ab,cb->ac", [(128, 32), (128, 32)]),          # one tile, one k-block, both K-major
    ("ab,cb->ac", [(128, 256), (128, 256)]),        # several k-blocks (stage wrap)
    ("ab,cb->ac", [(512, 512), (384, 512)]),        # several tiles per CTA? (12 tiles)
    ("ab,bc->ac", [(256, 128), (128, 256)]),        # B MN-major
    ("ba,bc->ac", [(128, 256), (128, 256)]),        # both MN-major
    ("ba,cb->ac", [(96, 256), (256, 96)]),          # A MN-major, B K-major
    ("ab,bc->ac", [(129, 67), (67, 131)]),          # generic layouts, ragged
    ("ab,bc->ca", [(200, 100), (100, 300)]),        # transposed output
    ("br,prs->bps", [(4096, 512), (2, 512, 512)]),  # cfg 3 step
    ("bij,bjk->bik", [(6, 150, 70), (6, 70, 140)]), # batched
    ("ab,cb->ac", [(4096, 4096), (4096, 4096)]),    # many tiles per CTA
]
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for idx, (einstr, shapes) in enumerate(CASES):
    if only is not None and idx != only:
        continue
    rng = np.random.default_rng(idx)
    ops = [rng.uniform(0.1, 1.0, s).astype(np.float32) for s in shapes]
    lst = planner.contract_path(einstr, *shapes)[1]
    plan = engine.Plan(lst, shapes, np.float32)
    print(idx, einstr, shapes, plan.launch_kinds, end=" ... ", flush=True)
    t0 = time.time()
    T, c, fl = plan.execute_host(ops)
    want = np.einsum(einstr, *[o.astype(np.float64) for o in ops])
    got = T.astype(np.float64) * np.exp(c)
    err = np.abs(got / want - 1).max()
    print(f"max rel err {err:.3e}  ({time.time()-t0:.2f}s)", flush=True)
    if idx in (8, 10):
        dev = [torch.as_tensor(o).cuda() for o in ops]
        for _ in range(3):
            plan.execute_device(dev)
        _, _, ms = plan.execute_device(dev, timed=True)
        k = plan.launch_kinds.index("gemm")
        print(f"    gemm kernel {ms[k]*1e3:.1f} us -> {plan.flops/ms[k]/1e9:.1f} TFLOP/s (useful, = x3 on the tensor pipe)", flush=True)
