#!/usr/bin/env python
"""Benchmark of the TN.contract() hot path on B200 (contract in the task brief).

Workload (BASELINE.json configs[1], SURVEY.md section 8d cfg 2): an order-101
copy-node hub with 100 attached (2, B) float64 tensors, B = 2^20 batched
evaluations -- einsum "ab,ab,...,ab->ab", 99 pairwise Hadamard steps each
followed by ``stabilize`` in the reference; ONE fused n-ary kernel here.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  ``value`` = contractions/s with the operands
resident in HBM; ``e2e`` = the same through the public API
(``TN.make_contract_fun`` -> ``contract`` -> ``ctn_plan_execute_host``) with
pinned HOST buffers, H2D/D2H inside the timed region.  N > 1: one process per
GPU, the batch is sharded (weak scaling: B evaluations per GPU), the shards'
(mantissa, log-scale) pairs are merged to one global scale after a 24-byte NCCL all-gather.
"""
import argparse
import json
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="evaluations per GPU (default 2^20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def ncu_traffic_bytes():
    """DRAM bytes per chain-kernel launch from the committed `ncu --set full`
    capture of this same command (profiles/r1_cfg2_chain_kernel_raw.csv)."""
    import csv

    path = os.path.join(ROOT, "profiles", "r1_cfg2_chain_kernel_raw.csv")
    try:
        with open(path) as f:
            rows = list(csv.reader(f))
        hdr, units, vals = rows[0], rows[1], rows[2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        total = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            total += float(vals[i].replace(",", "")) * scale[units[i]]
        return total
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def build_network(TN, arrays):
    """cfg 2: hub `a` first, batch hub `b`, then the dense (2, B) nodes."""
    tn = TN()
    hub = tn.add_copy_node(len(arrays) + 1)
    hub_b = tn.add_copy_node(len(arrays) + 1)
    for i, arr in enumerate(arrays):
        node = tn.add_dense_node(arr)
        tn.connect_nodes(hub, node, i, 0)
        tn.connect_nodes(hub_b, node, i, 1)
    return tn


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


def cpu_reference_time(arrays, contraction_list, reps):
    """The oracle port (numpy restatement of _core_contract + stabilize) on the
    host cores; best of `reps`."""
    from oracle import contract_oracle as oracle

    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.core_contract(arrays, contraction_list)
        best = min(best, time.perf_counter() - t0)
    return best


def host_threads():
    try:
        from threadpoolctl import threadpool_info

        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference algorithm (numpy port, oracle/) on the host."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from contractn_b200 import planner

    batch = args.batch
    # ~3 s per full-size contraction on 8 cores: shrink the sample if K+W is large
    total = args.steps + args.warmup
    sample = batch
    while sample > 1 << 14 and total * 3.0 * sample / BATCH > 150.0:
        sample //= 2
    arrays = [a for _, a in host_operands(sample, 1002, pinned=False)]
    einstr = ",".join(["ab"] * N_VEC) + "->ab"
    lst = planner.contract_path(einstr, *[a.shape for a in arrays])[1]
    from oracle import contract_oracle as oracle

    for _ in range(args.warmup):
        oracle.core_contract(arrays, lst)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.core_contract(arrays, lst)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = (sample / batch) / dt
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": "contractions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2 hub-101 x (2,B) f64", "batch_per_gpu": batch, "n_operands": N_VEC,
                   "planner": planner.planner_name()},
        "cpu_baseline": {"value": value, "unit": "contractions/s", "cores": host_threads(), "kind": "port",
                         "sample": f"batch {sample} of {batch} per step, scaled linearly; numpy einsum steps are single-threaded"},
        "e2e": {"value": value, "unit": "contractions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from contractn_b200 import TN, engine, planner

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
    tn = build_network(TN, arrays)
    einstr = tn.einsum_str
    contract_fun = tn.make_contract_fun(optimize="auto", split_format=True)
    shapes = tuple(a.shape for a in arrays)
    lst = planner.cached_contract_list(einstr, shapes, optimize="auto")
    plan = engine.get_plan(lst, shapes, np.float64, strides=tuple(engine._elem_strides(a) for a in arrays))
    dev_ops = [t.to(dev, non_blocking=True) for t, _ in ops]
    torch.cuda.synchronize()

    from contractn_b200 import parallel

    def step_device():
        out, tail = plan.execute_device(dev_ops)
        if world > 1:
            out, tail = parallel.merge_batch_shards(out, tail)
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
    for _ in range(max(3, args.warmup)):
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
            "config": {"workload": "cfg2 hub-101 x (2,B) f64", "batch_per_gpu": batch, "n_operands": N_VEC,
                       "einsum": "ab x100 -> ab", "pairwise_steps": len(lst), "planner": planner.planner_name(),
                       "l2": "inputs (1.68 GB) exceed L2 (126 MB); no flush needed",
                       "parallelism": f"batch-shard x{world}" if world > 1 else "single GPU"},
            "path_tflops": plan.flops * value / world * world / 1e12,
            "evaluations_per_s": value * batch,
            "kernels": {kind: float(per_kernel[i]) for i, kind in enumerate(kinds)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": ncu_traffic_bytes(),
                         "kernel": "chain_kernel<double,true,true>",
                         "algorithmic_bytes": algo_bytes, "kernel_ms": chain_ms, "peak_source": peak_src},
            "e2e": {"value": e2e_value, "unit": "contractions/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                    "api": "TN.make_contract_fun()(params, inputs) on pinned numpy arrays"},
            # chain + materialize + state-init kernel per step (+ shard-stats and merge kernels when sharded)
            "gpu_launches": (plan.num_launches + 1 + (2 if world > 1 else 0)) * args.steps,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            sec = cpu_reference_time(arrays, lst, reps=2)
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "contractions/s", "cores": host_threads(),
                                    "kind": "port", "sample": "full cfg2 workload, best of 2 contractions",
                                    "seconds_per_contraction": sec}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
