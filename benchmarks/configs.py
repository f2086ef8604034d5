"""All five BASELINE.json configurations on one B200 (device-resident timing),
at the tiers of SURVEY.md section 8d, with the CPU oracle timed beside them
where it finishes in seconds.  One JSON line per configuration.

  python benchmarks/configs.py [cfg1 cfg2 cfg3 cfg3s cfg4s cfg4m cfg5 ...] [--no-cpu]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contractn_b200 import TN, engine, planner  # noqa: E402


# network builders and sweep paths live in the package
from contractn_b200.networks import chain_tn, hub_tn, mps_tn, cp_tn, tucker_tn, peps_tn  # noqa: E402,F401


CONFIGS = {
    "cfg1": ("README chain: vector + 1000 3x3, f64", lambda: chain_tn(1000), True),
    "cfg2": ("hub-101 x (2, 2^20) f64", lambda: hub_tn(100, 1 << 20), True),
    "cfg2_f32": ("hub-101 x (2, 2^20) f32", lambda: hub_tn(100, 1 << 20, np.float32), False),
    "cfg3s": ("batched MPS 64 sites, bond 128, batch 1024, f32 (parity tier)", lambda: mps_tn(64, 128, 2, 1024), True),
    "cfg3": ("batched MPS 512 sites, bond 512, batch 4096, f32 (literal)", lambda: mps_tn(512, 512, 2, 4096), False),
    "cfg4s": ("CP I=256 R=256 f64 (tier S)", lambda: cp_tn(256, 256), True),
    "cfg4m": ("CP I=1024 R=1024 f64 (tier M)", lambda: cp_tn(1024, 1024), False),
    "cfg4t": ("Tucker I=512 R=256 f64", lambda: tucker_tn(512, 256), True),
    "cfg5s": ("PEPS 6x6 D=4 f64 (parity tier)", lambda: peps_tn(6, 4), True),
    "cfg5": ("PEPS 8x8 D=8 f64 (parity tier of SURVEY 8d)", lambda: peps_tn(8, 8), False),
}


def run(name, with_cpu):
    title, build, cpu_ok = CONFIGS[name]
    t0 = time.time()
    tn, optimize = build()
    einstr = tn.einsum_str
    params = tn.params
    shapes = tuple(tuple(p.shape) for p in params)
    dtype = params[0].dtype
    t_build = time.time() - t0
    t0 = time.time()
    lst = planner.cached_contract_list(einstr, shapes, optimize=optimize)
    t_plan = time.time() - t0
    plan = engine.Plan(lst, shapes, dtype)
    dev = [torch.as_tensor(p).cuda() for p in params]
    out, tail = plan.execute_device(dev)
    torch.cuda.synchronize()
    for _ in range(2):
        plan.execute_device(dev)
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out, tail = plan.execute_device(dev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    _, _, kms = plan.execute_device(dev, timed=True)
    by_kind = {}
    for kind, t in zip(plan.launch_kinds, kms):
        by_kind[kind] = by_kind.get(kind, 0.0) + t
    rec = {
        "config": name, "what": title, "dtype": str(dtype), "operands": len(params), "pairwise_steps": len(lst),
        "kernels": plan.num_launches, "path_flops": plan.flops, "step_bytes": plan.step_bytes,
        "ms": ms, "contractions_per_s": 1e3 / ms, "path_tflops": plan.flops / ms / 1e9,
        "GBps": plan.step_bytes / ms / 1e6, "ms_by_kind": by_kind,
        "workspace_MB": plan.workspace_bytes / 1e6, "host_build_s": t_build, "host_plan_s": t_plan,
        "log_scale": float(tail[0].item()),
    }
    if with_cpu and cpu_ok:
        from oracle import contract_oracle as oracle

        t0 = time.perf_counter()
        T0, c0 = oracle.core_contract(params, lst)
        rec["cpu_oracle_s"] = time.perf_counter() - t0
        rec["speedup_vs_cpu"] = rec["cpu_oracle_s"] * 1e3 / ms
        T = out.cpu().numpy().astype(np.float64)
        T0 = np.asarray(T0, dtype=np.float64)
        denom = np.abs(T0).max() or 1.0
        rec["max_rel_err_T"] = float(np.abs(T - T0).max() / denom)
        rec["abs_err_log_scale"] = abs(rec["log_scale"] - float(c0))
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg1", "cfg2", "cfg3s", "cfg4s", "cfg4t", "cfg5s"]
    for nm in names:
        run(nm, "--no-cpu" not in sys.argv)
