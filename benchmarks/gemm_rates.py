"""Device-resident rate of single GEMM-class steps (float64) vs the measured
cuBLAS DGEMM rate (profiles/r1_peaks.json).  python benchmarks/gemm_rates.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contractn_b200 import engine, planner  # noqa: E402

CASES = [
    ("ab,bc->ac", [(4096, 4096), (4096, 4096)]),
    ("ab,cb->ac", [(4096, 4096), (4096, 4096)]),
    ("ba,bc->ac", [(4096, 4096), (4096, 4096)]),
    ("br,prs->bps", [(4096, 512), (2, 512, 512)]),
    ("cab,da->bcd", [(256, 1024, 256), (256, 1024)]),
    ("dbc,eb->dce", [(256, 256, 256), (1024, 256)]),
    ("abcdefgh,efxy->abcdghxy", [(8,) * 8, (8, 8, 8, 8)]),
    ("abcdefg,gx->abcdefx", [(8,) * 7, (8, 8)]),
    ("ab,ab->", [(8192, 8192), (8192, 8192)]),
]


def main():
    f32 = "--f32" in sys.argv
    only = [a for a in sys.argv[1:] if a.isdigit()]
    peak = 756.5 / 3 if f32 else 35.45   # measured cuBLAS TF32 / 3 (3xTF32) or DGEMM
    rows = []
    for idx, (einstr, shapes) in enumerate(CASES):
        if only and str(idx) not in only:
            continue
        tdt, ndt = (torch.float32, np.float32) if f32 else (torch.float64, np.float64)
        ops = [torch.rand(s, dtype=tdt, device="cuda") for s in shapes]
        lst = planner.contract_path(einstr, *shapes)[1]
        plan = engine.Plan(lst, shapes, ndt)
        if "gemm" not in plan.launch_kinds:
            print(json.dumps({"einsum": einstr, "kinds": plan.launch_kinds, "note": "no gemm launch"}), flush=True)
            continue
        for _ in range(3):
            plan.execute_device(ops)
        torch.cuda.synchronize()
        best = None
        for _ in range(5):
            _, _, ms = plan.execute_device(ops, timed=True)
            best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
        k = plan.launch_kinds.index("gemm")
        gemm_ms = best[k] + (best[k + 1] if "gemm_reduce" in plan.launch_kinds else 0.0)
        tf = plan.flops / gemm_ms / 1e9
        gbs = plan.step_bytes / gemm_ms / 1e6
        rows.append({"einsum": einstr, "shapes": shapes, "kinds": plan.launch_kinds, "gemm_ms": gemm_ms,
                     "tflops": tf, "frac_of_peak": tf / peak, "peak_tflops": peak, "GBps": gbs})
        print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
