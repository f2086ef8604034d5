"""Bring-up / regression check of the tcgen05 3xTF32 GEMM kernel: progressively harder cases, each
printed as soon as it finishes (run under `timeout`).  For every case the split-format result is
compared with the float64 truth: max relative error of the mantissas and the error of the log-scale
(which is where an accumulator-truncation bias would show).

  python tools/t32_check.py [case index ...]        CTN_TF32_V1=1 selects the round-1 kernel
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contractn_b200 import engine, planner  # noqa: E402

CASES = [
    # (einsum, shapes, explicit path or None, timed?)
    ("ab,cb->ac", [(128, 32), (128, 32)], None, False),            # 0 one tile, one k-block, B K-major
    ("ab,cb->ac", [(128, 256), (128, 256)], None, False),          # 1 several k-blocks (stage wrap)
    ("ab,cb->ac", [(512, 512), (384, 512)], None, False),          # 2 several tiles
    ("ab,bc->ac", [(256, 128), (128, 256)], None, False),          # 3 B MN-major
    ("ba,bc->ac", [(128, 256), (128, 256)], None, False),          # 4 A M-fast (scalar coalesced loads)
    ("ba,cb->ac", [(96, 256), (256, 96)], None, False),            # 5 A M-fast, B K-major, ragged
    ("ab,bc->ac", [(129, 67), (67, 131)], None, False),            # 6 generic layouts, ragged
    ("ab,bc->ca", [(200, 100), (100, 300)], None, False),          # 7 transposed output
    ("br,prs->bps", [(4096, 512), (2, 512, 512)], None, True),     # 8 cfg 3 GEMM step alone
    ("bij,bjk->bik", [(6, 150, 70), (6, 70, 140)], None, False),   # 9 batched
    ("ab,cb->ac", [(4096, 4096), (4096, 4096)], None, True),       # 10 many tiles per CTA, K-major B (TMA)
    ("ab,bc->ac", [(4096, 4096), (4096, 4096)], None, True),       # 11 MN-major B (TMA, 128B_ATOM_32B)
    ("br,prs,bp->bs", [(4096, 512), (2, 512, 512), (4096, 2)], ((0, 1), (0, 1)), True),   # 12 cfg 3 fused site
    ("br,prs,bp->bs", [(1000, 96), (2, 96, 160), (1000, 2)], ((0, 1), (0, 1)), False),    # 13 fused, ragged
    ("br,prs,bp->bs", [(512, 64), (4, 64, 64), (512, 4)], ((0, 1), (0, 1)), False),       # 14 fused, phys dim 4
    ("ab,bc,cd->ad", [(4096, 512), (512, 1024), (1024, 512)], ((0, 1), (0, 1)), True),    # 15 two GEMMs: M-fast intermediate
    ("ab,bc,cd->ad", [(300, 100), (100, 260), (260, 90)], ((0, 1), (0, 1)), False),       # 16 same, ragged
    ("abk,kc->abc", [(40, 50, 64), (64, 72)], None, False),        # 17 two row modes
    ("akb,kc->abc", [(40, 64, 52), (64, 72)], None, False),        # 18 summed mode mid-tensor in A
    ("aklb,klc->abc", [(24, 8, 8, 20), (8, 8, 72)], None, False),  # 19 two summed modes
    ("ab,cb->ac", [(2048, 4096), (2048, 4096)], None, True),       # 20 long K
    ("br,prs,bp,qst,bq->bt", [(4096, 512), (2, 512, 512), (4096, 2), (2, 512, 512), (4096, 2)],
     ((0, 1), (0, 3), (0, 2), (0, 1)), True),                      # 21 two cfg 3 sites: M-fast intermediate
    ("ab,cb->ac", [(128, 32), (128, 32)], None, True),             # 22 one tile, one k-block: fixed cost of a launch
    ("ab,cb->ac", [(128 * 148, 512), (128, 512)], None, True),     # 23 one tile per CTA, 16 k-blocks
]


def run(idx):
    einstr, shapes, path, timed = CASES[idx]
    rng = np.random.default_rng(idx)
    ops = [rng.uniform(0.1, 1.0, s).astype(np.float32) for s in shapes]
    lst = planner.contract_path(einstr, *shapes, optimize=path if path is not None else "auto")[1]
    plan = engine.Plan(lst, shapes, np.float32)
    print(idx, einstr, shapes, plan.launch_kinds, end=" ... ", flush=True)
    t0 = time.time()
    T, c, fl = plan.execute_host(ops)
    if timed:  # big case: float64 reference on the GPU (the CPU einsum takes most of a minute)
        import torch

        want = torch.einsum(einstr, *[torch.as_tensor(o).cuda().double() for o in ops]).cpu().numpy()
    else:
        want = np.einsum(einstr, *[o.astype(np.float64) for o in ops], optimize=True)
    kappa = np.abs(want).mean()
    err_T = np.abs(T.astype(np.float64) / (want / kappa) - 1).max()
    err_c = c - np.log(kappa)
    print(f"T-hat max rel err {err_T:.2e}  log-scale err {err_c:+.2e}  flags {fl}  ({time.time() - t0:.2f}s)", flush=True)
    if timed:
        import torch

        dev = [torch.as_tensor(o).cuda() for o in ops]
        for _ in range(3):
            plan.execute_device(dev)
        best = None
        for _ in range(3):
            _, _, ms = plan.execute_device(dev, timed=True)
            best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
        for kind, t in zip(plan.launch_kinds, best):
            if kind == "gemm":
                print(f"    gemm kernel {t * 1e3:.1f} us", end="")
        print(f"   plan {plan.flops / sum(best) / 1e9:.1f} TFLOP/s useful (x3 on the tensor pipe)", flush=True)
    return err_T, err_c


if __name__ == "__main__":
    which = [int(a) for a in sys.argv[1:]] or range(len(CASES))
    bad = 0
    for i in which:
        eT, ec = run(i)
        if not (eT < 1e-5 and abs(ec) < 2e-6):
            bad += 1
            print("    ^^^ OUT OF TOLERANCE", flush=True)
    print("done, out of tolerance:", bad, flush=True)
