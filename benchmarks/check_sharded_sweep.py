"""Boundary-sharded sweep across GPUs (contractn_b200.sharded): parity against the float64 oracle and
against the unsharded one-GPU contraction, timing against the UNSHARDED contraction on one GPU, and a
per-segment time line (compute / scale merge / exchange) from CUDA events.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29541 benchmarks/check_sharded_sweep.py [L] [D] [--no-oracle]"""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contractn_b200 import engine, networks, parallel, planner, sharded

args = [a for a in sys.argv[1:] if not a.startswith("--")]
L = int(args[0]) if args else 8
D = int(args[1]) if len(args) > 1 else 8
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

tn, path = networks.peps_tn(L, D)
einstr = tn.einsum_str
ops = [torch.as_tensor(p).to(dev) for p in tn.params]
shapes = tuple(tuple(o.shape) for o in ops)
lst = planner.cached_contract_list(einstr, shapes, optimize=path)

sw = sharded.ShardedSweep(einstr, ops, optimize=path)
for _ in range(3):
    T, c = sw.run()
torch.cuda.synchronize()
times = []
for _ in range(10):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    T, c = sw.run()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t))
rec = {"world": world, "workload": f"closed {L}x{L} grid, D = {D}, f64, row-by-row sweep path",
       "exchange": "one kernel over peer memory (ctn_reshard_peer)" if sw._peer is not None else "merge kernel + NCCL all_to_all_single",
       "segments": len(sw.segments), "exchanges": sum(s.exchange is not None for s in sw.segments),
       "exchanged_MB_per_rank": sw.exchanged_bytes / 1e6,
       "seconds_best": min(times), "seconds_median": sorted(times)[len(times) // 2]}

# time line of one run: events around every segment's compute and exchange
marks = []
orig_seg, orig_merge = sw._run_segment, sharded.merge_batch_shards
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
def seg_timed(k, seg, o):
    a = ev(); r = orig_seg(k, seg, o); marks.append((k, "compute", a, ev())); return r
def merge_timed(*a, **k):
    x = ev(); r = orig_merge(*a, **k); marks.append((-1, "merge", x, ev())); return r
orig_a2a, orig_ag = dist.all_to_all_single, dist.all_gather_into_tensor
def a2a_timed(*a, **k):
    x = ev(); r = orig_a2a(*a, **k); marks.append((-1, "a2a", x, ev())); return r
def ag_timed(*a, **k):
    x = ev(); r = orig_ag(*a, **k); marks.append((-1, "gather", x, ev())); return r
orig_rs = parallel.reshard_peer
def rs_timed(*a, **k):
    x = ev(); r = orig_rs(*a, **k); marks.append((-1, "reshard", x, ev())); return r
parallel.reshard_peer = rs_timed
sw._run_segment = seg_timed; sharded.merge_batch_shards = merge_timed
dist.all_to_all_single = a2a_timed; dist.all_gather_into_tensor = ag_timed
e0 = ev(); sw.run(); e1 = ev(); torch.cuda.synchronize()
sw._run_segment = orig_seg; sharded.merge_batch_shards = orig_merge
dist.all_to_all_single = orig_a2a; dist.all_gather_into_tensor = orig_ag; parallel.reshard_peer = orig_rs
tot = lambda kind: sum(a.elapsed_time(b) for _, k, a, b in marks if k == kind)
rec["timeline_ms"] = {"total": e0.elapsed_time(e1), "segments_compute": tot("compute"), "scale_merges": tot("merge"),
                      "all_to_all": tot("a2a"), "all_gather": tot("gather"), "fused_reshard_kernels": tot("reshard"),
                      "per_fused_reshard": [round(a.elapsed_time(b), 3) for _, k, a, b in marks if k == "reshard"],
                      "per_segment_compute": [round(a.elapsed_time(b), 3) for _, k, a, b in marks if k == "compute"],
                      "per_all_to_all": [round(a.elapsed_time(b), 3) for _, k, a, b in marks if k == "a2a"]}
ks = 3 if len(sw.segments) > 3 else 0
hit = sw._bound[ks]
_, _, ms = hit[2].execute_device(hit[1].tensors, out=hit[1].out, timed=True)
rec["segment_kernels"] = {"segment": ks, "einsum": sw.segments[ks].einstr(), "ms": [round(m, 4) for m in ms],
                          "plan": hit[2].describe().splitlines()}

if sw._peer is not None and "--exchange-bench" in sys.argv:
    # the exchange kernel alone, back to back on the largest slab (no compute in between: no skew from it)
    ka = max((k for k, s in enumerate(sw.segments) if s.exchange == "a2a"), key=lambda k: sw._bound[k][1].out.numel())
    hit = sw._bound[ka]
    hit[1].run()
    for _ in range(3):
        parallel.reshard_peer(hit[1].out, hit[2], sw._peer, 0)
    sweep = []
    for ctas in (148, 296, 444, 592, 888, 1184):
        os.environ["CTN_RESHARD_CTAS"] = str(ctas)
        parallel.reshard_peer(hit[1].out, hit[2], sw._peer, 0)
        torch.cuda.synchronize(); dist.barrier()
        a = ev()
        for _ in range(20):
            parallel.reshard_peer(hit[1].out, hit[2], sw._peer, 0)
        b = ev(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 20], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nbytes = hit[1].out.numel() * 8
        sweep.append({"ctas": ctas, "ms": float(t), "remote_GBps_per_rank": nbytes * (world - 1) / world / float(t) / 1e6})
    del os.environ["CTN_RESHARD_CTAS"]
    rec["exchange_alone"] = {"slab_MB": nbytes / 1e6, "remote_MB": nbytes * (world - 1) / world / 1e6, "sweep": sweep}
if rank == 0:
    plan = engine.get_plan(lst, shapes, np.float64, device=dev.index)
    b = plan.bind(ops)
    b.run(); torch.cuda.synchronize()
    one = []
    for _ in range(5):
        t0 = time.perf_counter(); Tu, tail = b.run(); torch.cuda.synchronize(); one.append(time.perf_counter() - t0)
    cu = float(tail[0].item())
    rec.update({"seconds_unsharded_1gpu": min(one), "speedup_vs_unsharded_1gpu": min(one) / min(times),
                "path_flops": planner.path_flops(lst, einstr, shapes),
                "aggregate_path_tflops": planner.path_flops(lst, einstr, shapes) / min(times) / 1e12,
                "max_rel_err_T_vs_unsharded": float((T.double() - Tu.double()).abs().max() / Tu.double().abs().max()),
                "abs_err_log_scale_vs_unsharded": abs(float(c.reshape(-1)[0]) - cu)})
    if "--no-oracle" not in sys.argv:
        from oracle import contract_oracle as oracle
        t0 = time.perf_counter()
        To, co = oracle.contract(lst, *[np.asarray(p) for p in tn.params], split_format=True)
        rec.update({"oracle_seconds": time.perf_counter() - t0,
                    "max_rel_err_T_vs_oracle": float(np.abs(T.cpu().numpy() - np.asarray(To)).max() / np.abs(To).max()),
                    "abs_err_log_scale_vs_oracle": abs(float(c.reshape(-1)[0]) - float(co))})
    print(json.dumps(rec))
    os.makedirs("gpurun_out", exist_ok=True)
    tag = "" if sw._peer is not None or world == 1 else "_nccl"
    with open(f"gpurun_out/sharded_sweep_{L}x{L}_D{D}_{world}gpu{tag}.json", "w") as f:
        json.dump(rec, f, indent=1)
    assert rec["max_rel_err_T_vs_unsharded"] < 1e-11 and rec["abs_err_log_scale_vs_unsharded"] < 1e-9
if world > 1:
    dist.barrier(); dist.destroy_process_group()
