#!/usr/bin/env python
"""Benchmark of the TN.contract() hot path on B200 (contract in the task brief).

Headline workload (BASELINE.json configs[1], SURVEY.md section 8d cfg 2): an order-101
copy-node hub with 100 attached (2, B) float64 tensors, B = 2^20 batched evaluations --
einsum "ab,ab,...,ab->ab", 99 pairwise Hadamard steps each followed by ``stabilize`` in
the reference; ONE fused n-ary kernel here.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--configs LIST|none]

One JSON line on stdout (rank 0).  ``value`` = contractions/s with the operands resident
in HBM; ``e2e`` = the same through the public API (``TN.make_contract_fun`` ->
``contract`` -> ``ctn_plan_execute_host``) with pinned HOST buffers, H2D/D2H inside the
timed region.  ``configs`` (N = 1 only) carries one record per other BASELINE
configuration -- cfg 1 (README chain), cfg 3 literal (batched MPS, float32, the tcgen05
kernel), cfg 4 tier M CP + Tucker (FP64 DMMA), cfg 5 at its parity tier 8x8 D=8 -- each
with its own roofline, end-to-end number, CPU baseline (the oracle port) and errors
against the float64 oracle on the same inputs.  ``peaks`` are the denominators: HBM from
MEASURED_PEAKS.json, FP64 and TF32 GEMM measured HERE with cuBLAS (torch.matmul 8192^3).
N > 1: one process per GPU, the batch is sharded (weak scaling: B evaluations per GPU),
the shards' (mantissa, log-scale) pairs are merged to one global scale on the device.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_VEC = 100
BATCH = 1 << 20
METRIC = "contractions/s (cfg2: order-101 copy hub x 100 (2,B) f64, B=2^20 batched evaluations per GPU)"
DEFAULT_CONFIGS = "cfg1,cfg3,cfg4m,cfg4t,cfg5"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="evaluations per GPU (default 2^20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default=DEFAULT_CONFIGS,
                    help="other BASELINE configurations to measure at N=1 (comma list or 'none')")
    return ap.parse_args()


def headline_config(batch, world):
    """The `config` object of the JSON line: identical for both arms by construction."""
    from contractn_b200 import planner

    return {"workload": "cfg2 hub-101 x (2,B) f64", "batch_per_gpu": batch, "n_operands": N_VEC,
            "einsum": "ab x100 -> ab", "pairwise_steps": N_VEC - 1, "planner": planner.planner_name(),
            "l2": "inputs (1.68 GB) exceed L2 (126 MB); no flush needed",
            "parallelism": f"batch-shard x{world}" if world > 1 else "single GPU"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cublas_peaks():
    """FP64 and TF32 GEMM throughput of this GPU, measured now with cuBLAS (torch.matmul, 8192^3,
    best of 5 after warm-up): the denominators of the tensor-bound rooflines below."""
    import torch

    out = {}
    n = 8192
    for name, dt, tf32 in (("f64_tflops", torch.float64, False), ("tf32_tflops", torch.float32, True)):
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        a = torch.rand(n, n, dtype=dt, device="cuda")
        b = torch.rand(n, n, dtype=dt, device="cuda")
        c = torch.empty(n, n, dtype=dt, device="cuda")
        for _ in range(2):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b, c
    torch.cuda.empty_cache()
    out["how"] = "torch.matmul 8192^3 (cuBLAS), best of 5, CUDA events, measured in this run"
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def host_operands(batch, seed, pinned):
    rng = np.random.default_rng(seed)
    out = []
    if pinned:
        import torch

        for _ in range(N_VEC):
            t = torch.empty((2, batch), dtype=torch.float64, pin_memory=True)
            a = t.numpy()
            a[...] = rng.uniform(0.9, 1.1, (2, batch))
            out.append((t, a))
        return out
    return [(None, rng.uniform(0.9, 1.1, (2, batch))) for _ in range(N_VEC)]


def host_threads():
    try:
        from threadpoolctl import threadpool_info

        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


def oracle_contract_fun(einstr, shapes, optimize):
    """The reference's ``contract_fun`` (ctn.py:378-385) restated on the oracle: arg packing is the
    identity here, then ``contract`` = cached path lookup + ``_core_contract`` + split output."""
    from contractn_b200 import planner
    from oracle import contract_oracle as oracle

    def fun(arrays):
        lst = planner.cached_contract_list(einstr, shapes, optimize=optimize)  # lru, like einsum.py:313
        return oracle.core_contract(arrays, lst)

    return fun


def run_reference(args):
    """--impl reference: the reference algorithm (numpy port, oracle/) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    batch = args.batch
    # ~1 s per full-size contraction on the host: shrink the sample if K + W is large
    total = args.steps + args.warmup
    sample = batch
    while sample > 1 << 14 and total * 1.2 * sample / BATCH > 150.0:
        sample //= 2
    arrays = [a for _, a in host_operands(sample, 1002, pinned=False)]
    einstr = ",".join(["ab"] * N_VEC) + "->ab"
    fun = oracle_contract_fun(einstr, tuple(a.shape for a in arrays), "auto")
    for _ in range(args.warmup):
        fun(arrays)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fun(arrays)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = (sample / batch) / dt
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": "contractions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": headline_config(batch, world),
        "cpu_baseline": {"value": value, "unit": "contractions/s", "cores": host_threads(), "kind": "port",
                         "sample": f"batch {sample} of {batch} per step, scaled linearly; contract_fun-equivalent "
                                   "(cached path lookup + executor); numpy einsum steps are single-threaded"},
        "e2e": {"value": value, "unit": "contractions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# the other BASELINE configurations (N = 1)
# ---------------------------------------------------------------------------
def _time_device(plan, dev_ops, steps):
    import torch

    bound = plan.bind(dev_ops)
    for _ in range(3):
        bound.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out, tail = bound.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    per = np.zeros(plan.num_launches)
    reps = 3
    for _ in range(reps):
        _, _, kms = plan.execute_device(dev_ops, timed=True)
        per += np.array(kms)
    return ms, per / reps, out, tail


def run_config(name, peaks, hbm_peak, with_cpu):
    """One record of the `configs` array."""
    import torch

    from contractn_b200 import engine, networks, planner
    from oracle import contract_oracle as oracle

    spec = {
        "cfg1": ("cfg1 README chain: vector + 1000 x (3,3), f64", lambda: networks.chain_tn(1000), 20),
        "cfg3": ("cfg3 batched MPS: 512 sites, bond 512, phys 2, batch 4096, f32, explicit sweep path (literal size)",
                 lambda: networks.mps_tn(512, 512, 2, 4096), 5),
        "cfg4m": ("cfg4 CP tier M: 3 factors (1024, 1024) on an order-3 hub, f64 (8.6 GB output)",
                  lambda: networks.cp_tn(1024, 1024), 3),
        "cfg4t": ("cfg4 Tucker: core (256,256,256), factors (512, 256), f64", lambda: networks.tucker_tn(512, 256), 5),
        "cfg5": ("cfg5 PEPS parity tier: closed 8x8 grid, D = 8, f64, row-by-row sweep path (6.5e11 FLOP)",
                 lambda: networks.peps_tn(8, 8), 5),
    }[name]
    title, build, steps = spec
    tn, optimize = build()
    einstr = tn.einsum_str
    params = tn.params
    shapes = tuple(tuple(p.shape) for p in params)
    dtype = params[0].dtype
    lst = planner.cached_contract_list(einstr, shapes, optimize=optimize)
    plan = engine.get_plan(lst, shapes, dtype, strides=tuple(engine._elem_strides(p) for p in params),
                           device=torch.cuda.current_device())
    dev_ops = [torch.as_tensor(p).cuda() for p in params]
    ms, per, out, tail = _time_device(plan, dev_ops, steps)
    kinds = plan.launch_kinds
    by_kind = {}
    for k, t in zip(kinds, per):
        by_kind[k] = by_kind.get(k, 0.0) + float(t)
    rec = {"name": name, "workload": title, "dtype": "f32" if dtype == np.float32 else "f64",
           "pairwise_steps": len(lst), "kernels_per_contraction": plan.num_launches,
           "ms_per_contraction": ms, "contractions_per_s": 1e3 / ms, "path_flops": plan.flops,
           "path_tflops": plan.flops / ms / 1e9, "algorithmic_GBps": plan.step_bytes / ms / 1e6,
           "ms_by_kernel_class": by_kind, "workspace_MB": plan.workspace_bytes / 1e6}

    # ---- roofline of the dominant kernel class ----
    if name == "cfg1":
        rec["roofline"] = {"bound": "latency", "kernel": "interp_kernel (one launch, 1000 dependent steps)",
                           "us_per_dependent_step": ms * 1e3 / len(lst), "achieved": None, "peak": None, "frac": None,
                           "unit": None, "traffic": None}
    else:
        gemm_ms = by_kind.get("gemm", 0.0)
        # FLOPs of the GEMM-class launches = the plan's path FLOPs minus the (few) non-GEMM steps; for these
        # configurations the GEMM steps carry > 99.9 % of the path FLOPs, so the plan total is used
        if dtype == np.float32:
            peak, unit_name = peaks["tf32_tflops"] / 3.0, "gemm_tf32x3_kernel (tcgen05 kind::tf32, 3 MMAs per product)"
            src = "cuBLAS TF32 8192^3 measured in this run, divided by 3 (error-compensated split)"
        else:
            peak, unit_name = peaks["f64_tflops"], "gemm_f64_kernel / skinny_dmma_kernel (DMMA)"
            src = "cuBLAS DGEMM 8192^3 measured in this run"
        ach = plan.flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        rec["roofline"] = {"bound": "tensor", "kernel": unit_name, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                           "frac": ach / peak if ach else None, "traffic": None, "kernel_ms": gemm_ms,
                           "share_of_step": gemm_ms / float(per.sum()) if per.sum() > 0 else None, "peak_source": src}

    # ---- end to end through the public API on host arrays ----
    try:
        contract_fun = tn.make_contract_fun(optimize=optimize, split_format=True)
        contract_fun(params, ())
        n_e2e = 1 if name == "cfg4m" else 3
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            T_h, c_h = contract_fun(params, ())
        e2e_s = (time.perf_counter() - t0) / n_e2e
        rec["e2e"] = {"value": 1.0 / e2e_s, "unit": "contractions/s", "ms_per_step": e2e_s * 1e3,
                      "h2d_bytes_per_step": int(sum(p.nbytes for p in params)),
                      "d2h_bytes_per_step": int(np.asarray(T_h).nbytes + 16), "steps": n_e2e,
                      "api": "TN.make_contract_fun()(params, inputs) on host numpy arrays"}
        del T_h
    except MemoryError as exc:  # the 8.6 GB host result of cfg4m on a small-memory box
        rec["e2e"] = {"value": None, "error": str(exc)}

    # ---- errors against the float64 oracle on the same inputs, and the CPU baseline ----
    c_gpu = float(tail[0].item())
    if name == "cfg4m":
        # the oracle needs minutes and 17 GB here: check 4000 sampled entries against the defining sum
        rng = np.random.default_rng(4)
        idx = rng.integers(0, 1024, size=(4000, 3))
        got = out[idx[:, 0], idx[:, 1], idx[:, 2]].cpu().numpy() * math.exp(c_gpu)
        f1, f2, f3 = params
        want = np.einsum("na,na,na->n", f1[idx[:, 0]], f2[idx[:, 1]], f3[idx[:, 2]])
        rec["errors"] = {"max_rel_err_sampled_entries": float(np.abs(got / want - 1).max()), "samples": 4000,
                         "mean_abs_T_minus_1": abs(float(out.abs().mean().item()) - 1.0),
                         "vs": "defining sum on sampled entries (float64) + canonical form mean|T| = 1"}
        if with_cpu:
            tn_s, opt_s = networks.cp_tn(256, 256)
            ps = tn_s.params
            fun = oracle_contract_fun(tn_s.einsum_str, tuple(p.shape for p in ps), opt_s)
            fun(ps)
            t0 = time.perf_counter()
            fun(ps)
            sec = time.perf_counter() - t0
            scale = (1024 / 256) ** 3 * (1024 / 256)  # I^3 * R work
            rec["cpu_baseline"] = {"value": 1.0 / (sec * scale), "unit": "contractions/s", "cores": host_threads(),
                                   "kind": "port", "sample": f"tier S (I = R = 256) in {sec:.3f} s, scaled by I^3 R = {scale:.0f}x"}
    else:
        ops64 = [np.asarray(p, dtype=np.float64) for p in params]
        t0 = time.perf_counter()
        T0, c0 = oracle.core_contract(ops64, lst)
        sec64 = time.perf_counter() - t0
        T = out.cpu().numpy().astype(np.float64)
        T0 = np.asarray(T0, dtype=np.float64)
        rec["errors"] = {"max_rel_err_T": float(np.abs(T - T0).max() / (np.abs(T0).max() or 1.0)),
                         "abs_err_log_scale": abs(c_gpu - float(c0)), "vs": "oracle in float64 on the same inputs"}
        if with_cpu:
            if dtype == np.float64:
                sec, how = sec64, "full workload, one contraction (the float64 oracle run used for the errors)"
            else:
                fun = oracle_contract_fun(einstr, shapes, optimize)
                t0 = time.perf_counter()
                fun(params)
                sec, how = time.perf_counter() - t0, "full workload in float32 (the reference's arithmetic), one contraction"
            rec["cpu_baseline"] = {"value": 1.0 / sec, "unit": "contractions/s", "cores": host_threads(), "kind": "port",
                                   "sample": how, "seconds_per_contraction": sec}
    del dev_ops, out
    torch.cuda.empty_cache()
    return rec


def run_multi_gpu_records(world, rank, dev):
    """N > 1 only: the two OTHER ways the path shards, measured on this job's GPUs.

    * ``slab``  -- cfg 4 CP reconstruction with the first open index split across the ranks (output-slab
      sharding: no tensor data is exchanged, one 24-byte scale exchange).  Tier M (I = R = 1024) at N = 2, 4;
      the literal tier L (I = 4096, R = 1024: 550 GB of output, 68.7 GB per GPU) at N = 8.
    * ``sliced`` -- cfg 5 at its parity tier (closed 8x8 grid, D = 8) with the most expensive bond sliced
      (8 slices round-robin over the ranks, log-sum-exp merge of the slice sums), timed against the UNSLICED
      contraction on one GPU: a bond lives in the boundary for only L of the L^2 sweep steps, so slicing
      multiplies the total work -- the ratio reported is the honest one.
    """
    import torch
    import torch.distributed as dist

    from contractn_b200 import engine, networks, parallel, planner

    def timed(fn, reps=2, free=True):
        best, res = None, None
        for _ in range(reps):
            res = None  # (release the previous result first: a tier-L slab is 68.7 GB)
            if free:  # tier L needs the memory back; otherwise the caching allocator keeps the blocks (a cudaMalloc
                torch.cuda.empty_cache()  # of gigabytes inside the timed region is not part of the contraction)
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = fn()
            torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item()) if best is None else min(best, float(t.item()))
        return best, res

    out = {}
    # ---------------- slab sharding (cfg 4) ----------------
    try:
        I, R = (4096, 1024) if world >= 8 else (1024, 1024)
        tn, opt = networks.cp_tn(I, R)
        einstr = tn.einsum_str
        fs = [torch.as_tensor(p).to(dev) for p in tn.params]
        sec, (T, c, (lo, hi)) = timed(lambda: parallel.contract_output_slabs(einstr, *fs, optimize=opt),
                                      reps=2 if I == 4096 else 4, free=I == 4096)
        rng = np.random.default_rng(40 + rank)
        idx = np.stack([rng.integers(lo, hi, 2000), rng.integers(0, I, 2000), rng.integers(0, I, 2000)], axis=1)
        terms = einstr.split("->")[0].split(",")
        got = T[idx[:, 0] - lo, idx[:, 1], idx[:, 2]].cpu().numpy() * math.exp(float(c[0]))
        # the defining sum (entry (i, j, k) depends on rows i, j, k of the factors in output-index order)
        outl = einstr.split("->")[1]
        facs = [tn.params[[t[0] for t in terms].index(l)] for l in outl]
        want = np.einsum("na,na,na->n", facs[0][idx[:, 0]], facs[1][idx[:, 1]], facs[2][idx[:, 2]])
        err = torch.tensor([float(np.abs(got / want - 1).max())], dtype=torch.float64, device=dev)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        flops = 2.0 * I * I * I * R + I * I * R
        rec = {"workload": f"cfg4 CP I={I} R={R} f64, output slabs over index '{outl[0]}' on {world} GPUs",
               "tier": "L (literal, 550 GB output)" if I == 4096 else "M", "seconds": sec,
               "aggregate_path_tflops": flops / sec / 1e12, "slab_GB_per_gpu": T.numel() * 8 / 1e9,
               "max_rel_err_sampled_entries": float(err.item()), "samples_per_rank": 2000}
        del T
        torch.cuda.empty_cache()
        if I == 1024:
            dist.barrier()
            if rank == 0:
                lst = planner.cached_contract_list(einstr, tuple(tuple(f.shape) for f in fs), optimize=opt)
                plan = engine.get_plan(lst, tuple(tuple(f.shape) for f in fs), np.float64, device=dev.index)
                b = plan.bind(fs)
                b.run()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                b.run()
                torch.cuda.synchronize()
                one = time.perf_counter() - t0
                rec["seconds_unsharded_1gpu"] = one
                rec["speedup_vs_unsharded_1gpu"] = one / sec
                del b, plan
                engine.clear_plan_cache()
                torch.cuda.empty_cache()
            dist.barrier()
        out["slab"] = rec
    except Exception as exc:
        out["slab"] = {"error": f"{type(exc).__name__}: {exc}"}
    # ---------------- bond slicing (cfg 5) ----------------
    try:
        tn, path = networks.peps_tn(8, 8)
        einstr = tn.einsum_str
        ops = [torch.as_tensor(p).to(dev) for p in tn.params]
        shapes = tuple(tuple(o.shape) for o in ops)
        lst = planner.cached_contract_list(einstr, shapes, optimize=path)
        labels = parallel.choose_slice_labels(einstr, shapes, lst, 8)
        sec, (Ts, cs) = timed(lambda: parallel.contract_sliced(einstr, *ops, slice_labels=labels, optimize=path,
                                                                split_format=True))
        rec = {"workload": "cfg5 parity tier: closed 8x8 grid, D = 8, f64; 8 slices of one bond, round-robin over "
                           f"{world} GPUs", "seconds": sec}
        if rank == 0:
            plan = engine.get_plan(lst, shapes, np.float64, device=dev.index)
            b = plan.bind(ops)
            b.run()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, tail = b.run()
            torch.cuda.synchronize()
            one = time.perf_counter() - t0
            net = parallel.SlicedNetwork(einstr, shapes, labels)
            lst_s = planner.cached_contract_list(net.einstr, net.shapes, optimize=path)
            rec.update({"seconds_unsliced_1gpu": one, "speedup_vs_unsliced_1gpu": one / sec,
                        "work_multiplier_of_slicing": net.n_slices * planner.path_flops(lst_s, net.einstr, net.shapes)
                        / planner.path_flops(lst, einstr, shapes),
                        "abs_err_log_scale_vs_unsliced": abs(float(cs[0]) - float(tail[0].item()))})
        dist.barrier()
        out["sliced"] = rec
    except Exception as exc:
        out["sliced"] = {"error": f"{type(exc).__name__}: {exc}"}
    # ---------------- boundary-sharded sweep (cfg 5, work stays 1x) ----------------
    engine.clear_plan_cache()   # (the cached slab / sliced plans keep their workspaces: tier L holds 68.7 GB)
    torch.cuda.empty_cache()
    for L in ((9, 8) if world >= 4 else (8,)):   # (largest first: the peer-mapped receive arena is allocated once)
        key = "sharded_sweep" if L == 8 else f"sharded_sweep_{L}x{L}"
        try:
            out[key] = sharded_sweep_record(L, 8, world, rank, dev)
        except Exception as exc:
            out[key] = {"error": f"{type(exc).__name__}: {exc}"}
        engine.clear_plan_cache()
        torch.cuda.empty_cache()
    return out


def sharded_sweep_record(L, D, world, rank, dev):
    """cfg 5 (closed L x L grid, bond D, f64, row-by-row sweep) with the boundary tensor SHARDED over one of its
    legs (contractn_b200/sharded.py): every pairwise step of the path is executed once, 1/W of it per rank; the
    boundary is re-sharded by one kernel over peer memory when the sharded leg is about to be absorbed.  Timed
    against the UNSHARDED contraction on one GPU (same path, same plan as the N = 1 `cfg5` record)."""
    import torch
    import torch.distributed as dist

    from contractn_b200 import engine, networks, planner, sharded

    tn, path = networks.peps_tn(L, D)
    einstr = tn.einsum_str
    ops = [torch.as_tensor(p).to(dev) for p in tn.params]
    shapes = tuple(tuple(o.shape) for o in ops)
    lst = planner.cached_contract_list(einstr, shapes, optimize=path)
    sw = sharded.ShardedSweep(einstr, ops, optimize=path)
    for _ in range(2):
        T, c = sw.run()
    times = []
    for _ in range(5 if L == 8 else 3):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        T, c = sw.run()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    flops = planner.path_flops(lst, einstr, shapes)
    rec = {"workload": f"cfg5 tier: closed {L}x{L} grid, D = {D}, f64, row-by-row sweep; boundary sharded over a leg "
                       f"on {world} GPUs", "seconds": min(times), "timing": "CUDA events, max over ranks, best of "
                       f"{len(times)}", "work_multiplier": 1.0, "path_flops": flops,
           "aggregate_path_tflops": flops / min(times) / 1e12,
           "segments": len(sw.segments), "exchanges": sum(s.exchange is not None for s in sw.segments),
           "exchanged_MB_per_rank": sw.exchanged_bytes / 1e6,
           "exchange": "ctn_reshard_peer: scale merge + all-to-all stores over NVLink peer memory, one kernel"
                       if sw._peer is not None else "merge kernel + NCCL all_to_all_single"}
    c_mine = float(c.reshape(-1)[0].item())
    T_mine = T.detach().double().cpu()
    del sw
    if rank == 0:
        plan = engine.get_plan(lst, shapes, np.float64, device=dev.index)
        b = plan.bind(ops)
        b.run()
        torch.cuda.synchronize()
        one = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            Tu, tail = b.run()
            e1.record()
            torch.cuda.synchronize()
            one.append(e0.elapsed_time(e1) / 1e3)
        Tu = Tu.double().cpu()
        rec.update({"seconds_unsharded_1gpu": min(one), "speedup_vs_unsharded_1gpu": min(one) / min(times),
                    "max_rel_err_T_vs_unsharded": float((T_mine - Tu).abs().max() / Tu.abs().max()),
                    "abs_err_log_scale_vs_unsharded": abs(c_mine - float(tail[0].item()))})
        del b, plan
    dist.barrier()
    return rec


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from contractn_b200 import engine, networks, planner

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    assert engine.load_library().ctn_device_count() > 0, "no sm_100 device"

    batch = args.batch
    ops = host_operands(batch, 1002 + rank, pinned=True)
    arrays = [a for _, a in ops]
    tn, _ = networks.hub_tn(N_VEC, batch, arrays=arrays)
    einstr = tn.einsum_str
    contract_fun = tn.make_contract_fun(optimize="auto", split_format=True)
    shapes = tuple(a.shape for a in arrays)
    lst = planner.cached_contract_list(einstr, shapes, optimize="auto")
    plan = engine.get_plan(lst, shapes, np.float64, strides=tuple(engine._elem_strides(a) for a in arrays),
                           device=local)
    dev_ops = [t.to(dev, non_blocking=True) for t, _ in ops]
    torch.cuda.synchronize()

    from contractn_b200 import parallel

    bound = plan.bind(dev_ops)  # operand pointers resolved once: the timed step is one C call (+ the merge)

    def step_device():
        out, tail = bound.run()
        if world > 1:
            out, tail = parallel.merge_batch_shards(out, tail, plan=plan)
        return out, tail

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------
    # the clock sampler (one long-running nvidia-smi) is started BEFORE the warm-up and given time to
    # print its first sample: its start-up queries take driver locks and would otherwise stall the
    # first timed launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        t_wait = time.perf_counter()
        while sampler.proc is not None and not sampler.lines and time.perf_counter() - t_wait < 3.0:
            time.sleep(0.02)
    for _ in range(max(3, args.warmup) + (20 if world > 1 else 0)):  # (the merge path warms up its peer exchange)
        step_device()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        # ranks leave the host barrier up to ~1 ms apart; a device-side collective right before the start
        # event lines the streams up, so that skew is not charged to the first merge of the timed steps
        dist.all_reduce(torch.zeros(1, device=dev))
    ev0.record()
    for _ in range(args.steps):
        out, tail = step_device()
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    # per-kernel device times (CUDA events between the kernels, same stream)
    per_kernel = np.zeros(plan.num_launches)
    reps = max(3, min(args.steps, 10))
    for _ in range(reps):
        _, _, kms = plan.execute_device(dev_ops, timed=True)
        per_kernel += np.array(kms)
    per_kernel /= reps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world / (ms_step * 1e-3)

    # ---- end to end through the public API, host buffers ------------------
    e2e_steps = max(2, min(args.steps, 10))
    for _ in range(2):
        contract_fun(tn.params, ())
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        T, c = contract_fun(tn.params, ())
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world / float(t.item())
    clocks = sampler.stop() if rank == 0 else None  # sampled over the device-timed AND e2e-timed regions
    h2d = int(sum(a.nbytes for a in arrays))
    d2h = int(T.nbytes + 16)

    multi = None
    if world > 1 and args.configs != "none":
        del dev_ops, bound
        torch.cuda.empty_cache()
        multi = run_multi_gpu_records(world, rank, dev)
    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        kinds = plan.launch_kinds
        k = kinds.index("chain")
        algo_bytes = plan.step_bytes  # every operand once + the output once
        chain_ms = float(per_kernel[k])
        achieved = algo_bytes / (chain_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "contractions/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": headline_config(batch, world),
            "path_tflops": plan.flops * value / 1e12,
            "evaluations_per_s": value * batch,
            "kernels": {kind: float(per_kernel[i]) for i, kind in enumerate(kinds)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": None,
                         "kernel": "chain_kernel<double,true,true>",
                         "algorithmic_bytes": algo_bytes, "kernel_ms": chain_ms, "peak_source": peak_src,
                         "traffic_note": "dram bytes per launch from `ncu --set full`: profiles/ (not re-measured in this run)"},
            "e2e": {"value": e2e_value, "unit": "contractions/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                    "h2d_GBps_per_gpu": h2d / e2e_s / 1e9,
                    "api": "TN.make_contract_fun()(params, inputs) on pinned numpy arrays"},
            # chain + materialize + state-init kernel per step (+ the merge kernel when sharded)
            "gpu_launches": (plan.num_launches + 1 + (1 if world > 1 else 0)) * args.steps,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            fun = oracle_contract_fun(einstr, shapes, "auto")
            best = float("inf")
            for _ in range(2):
                t0 = time.perf_counter()
                fun(arrays)
                best = min(best, time.perf_counter() - t0)
            line["cpu_baseline"] = {"value": 1.0 / best, "unit": "contractions/s", "cores": host_threads(),
                                    "kind": "port", "sample": "full cfg2 workload, best of 2 contractions",
                                    "seconds_per_contraction": best}
        if multi is not None:
            line["multi_gpu"] = multi
        if world == 1 and args.configs != "none":
            del dev_ops, bound
            torch.cuda.empty_cache()
            peaks = cublas_peaks()
            peaks["hbm_gbs"] = hbm_peak
            line["peaks"] = peaks
            recs = []
            for nm in [c for c in args.configs.split(",") if c]:
                try:
                    recs.append(run_config(nm, peaks, hbm_peak, not args.no_cpu_baseline))
                except Exception as exc:  # one configuration must not take the headline down with it
                    recs.append({"name": nm, "error": f"{type(exc).__name__}: {exc}"})
            line["configs"] = recs
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
