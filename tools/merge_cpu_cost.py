import os, sys, time, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
from contractn_b200 import engine, parallel, planner
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, B = 100, 1 << 20
ops = [torch.rand(2, B, dtype=torch.float64, device=dev) * 0.2 + 0.9 for _ in range(n)]
einstr = ",".join(["ab"] * n) + "->ab"
lst = planner.contract_path(einstr, *[tuple(t.shape) for t in ops])[1]
plan = engine.get_plan(lst, tuple(tuple(t.shape) for t in ops), np.float64)
def step(merge):
    out, tail = plan.execute_device(ops)
    if merge: out, tail = parallel.merge_batch_shards(out, tail)
    return out, tail
for merge in (False, True):
    for _ in range(5): step(merge)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(50): step(merge)
    e1.record(); t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    if rank == 0: print(json.dumps({"merge": merge, "cpu_enqueue_us_per_step": t_cpu / 50 * 1e6, "gpu_us_per_step": e0.elapsed_time(e1) / 50 * 1e3}), flush=True)
# pieces of the merge on the CPU side
out, tail = plan.execute_device(ops); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): parallel.merge_batch_shards(out, tail)
t1 = time.perf_counter() - t0; torch.cuda.synchronize()
if rank == 0: print("merge alone cpu us", t1 / 50 * 1e6)
ws, table = parallel._shard_buffers(dev, world)
t0 = time.perf_counter()
for _ in range(50): dist.all_gather_into_tensor(table, ws[:3])
t1 = time.perf_counter() - t0; torch.cuda.synchronize()
if rank == 0: print("all_gather_into_tensor cpu us", t1 / 50 * 1e6)
dist.barrier(); dist.destroy_process_group()
