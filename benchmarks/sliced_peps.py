"""cfg 5 across GPUs: closed LxL PEPS grid, bond slicing over the heaviest
bonds, slices round-robin over ranks, NCCL log-sum-exp merge.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 benchmarks/sliced_peps.py [L] [D] [min_slices] [--check]
(or plain `python benchmarks/sliced_peps.py` for one GPU).  One JSON line on rank 0.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contractn_b200 import engine, parallel, planner  # noqa: E402
from benchmarks.configs import peps_tn  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    L = int(args[0]) if len(args) > 0 else 8
    D = int(args[1]) if len(args) > 1 else 8
    min_slices = int(args[2]) if len(args) > 2 else 64
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tn, path = peps_tn(L, D)
    einstr = tn.einsum_str
    params = tn.params
    shapes = tuple(p.shape for p in params)
    lst = planner.cached_contract_list(einstr, shapes, optimize=path)
    labels = parallel.choose_slice_labels(einstr, shapes, lst, min_slices)
    net = parallel.SlicedNetwork(einstr, shapes, labels)
    ops = [torch.as_tensor(p).to(dev) for p in params]
    sliced_lst = planner.cached_contract_list(net.einstr, net.shapes, optimize=path)
    flops_per_slice = planner.path_flops(sliced_lst, net.einstr, net.shapes)

    def run():
        return parallel.contract_sliced(einstr, *ops, slice_labels=labels, optimize=path, split_format=True)

    T, c = run()  # warm-up (plan creation, workspace)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        T, c = run()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    rec = {
        "workload": f"PEPS {L}x{L} D={D} f64, {net.n_slices} slices over labels {len(labels)} bonds",
        "n_gpus": world, "ms": float(ms.item()), "contractions_per_s": 1e3 / float(ms.item()),
        "path_flops_unsliced": planner.path_flops(lst, einstr, shapes),
        "path_flops_sliced_total": flops_per_slice * net.n_slices,
        "path_tflops": flops_per_slice * net.n_slices / float(ms.item()) / 1e9,
        "log_scale": float(c[0].item()), "mantissa": float(T.reshape(-1)[0].item()),
    }
    if "--check" in sys.argv and rank == 0:
        T0, c0 = engine.contract(einstr, *ops, optimize=path, split_format=True)
        rec["unsliced_log_scale"] = float(c0.item())
        rec["abs_err_log_scale"] = abs(rec["log_scale"] - rec["unsliced_log_scale"])
        rec["mantissa_unsliced"] = float(T0.reshape(-1)[0].item())
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
