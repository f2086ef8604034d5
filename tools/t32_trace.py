"""Time line of one CTA of the tcgen05 kernel (CTN_T32_DEBUG=16): per k-block, when each pipeline event
happened on SM 0 (cycles relative to the kernel's start).  python tools/t32_trace.py <t32_check case>"""
import ctypes, os, sys
import numpy as np
os.environ["CTN_T32_DEBUG"] = str(16 | int(os.environ.get("CTN_T32_EXTRA", "0")))  # extra timing-experiment flags
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from contractn_b200 import engine, planner
import t32_check
import torch

idx = int(sys.argv[1])
einstr, shapes, path, _ = t32_check.CASES[idx]
rng = np.random.default_rng(idx)
ops = [torch.as_tensor(rng.uniform(0.1, 1.0, s).astype(np.float32)).cuda() for s in shapes]
lst = planner.contract_path(einstr, *shapes, optimize=path if path is not None else "auto")[1]
plan = engine.Plan(lst, shapes, np.float32)
lib = engine.load_library()
E, I = 32, 1024
buf = (ctypes.c_longlong * (E * I))()
for _ in range(3):
    plan.execute_device(ops)
torch.cuda.synchronize()
lib.ctn_debug_t32_trace(buf, E * I)  # drop the warm-up records
_, _, ms = plan.execute_device(ops, timed=True)
torch.cuda.synchronize()
print("launch durations by CUDA events (us):", [round(x * 1e3, 1) for x in ms], plan.launch_kinds)
lib.ctn_debug_t32_trace(buf, E * I)
t = np.array(buf[:], dtype=np.int64).reshape(E, I)
names = {0: "start", 1: "A raw landed", 2: "A tmem free", 3: "FullA", 4: "issued", 5: "B landed", 6: "FullB",
         7: "mma: A ready", 8: "mma: B ready", 9: "mma issued", 10: "epi: acc full", 12: "epi: tile done",
         20: "loader setup done", 30: "all roles done", 31: "end"}
# (with several gemm launches in the plan the LAST launch's records survive)
t0 = t[0, 0]
cols = [21, 22, 23, 24, 4, 1, 2, 3, 5, 6, 7, 8, 9]
names.update({21: "i:A free", 22: "i:A copies", 23: "i:A arrive", 24: "i:B free", 13: "epi: ld0 | pass 1 done", 14: "epi: ld1 | alpha ready", 15: "epi: ld2 | chunk 0 computed",
              16: "epi: ld3 | chunk 0 stored", 17: "epi: table ready", 27: "block sums done", 28: "ticket done"})
n_it = int((t[9] > 0).sum())
print(f"k-blocks recorded: {n_it}; span start->end {t[31, 0] - t0} cycles")
print("it    " + " ".join(f"{names[c][:12]:>12}" for c in cols))
show = list(range(min(n_it, 40))) + ([] if n_it <= 44 else list(range(n_it - 4, n_it)))
for it in show:
    print(f"{it:4d}  " + " ".join(f"{int(t[c, it] - t0):12d}" if t[c, it] > 0 else " " * 12 for c in cols))
for c in (20, 17, 10, 13, 14, 15, 16, 12, 30, 27, 28, 31):
    for it in range(8):
        if t[c, it] > 0:
            print(f"   {names[c]:>30} [{it}] at {int(t[c, it] - t0)}")
if t[11, 1] > 0:  # inside the A converters: landed -> stage free -> FullA -> [pending row dot: enter, sums there, dot done, left]
    print("A converter, k-blocks 8..15: landed->free  ->FullA  ->enter dot  ->sums there  ->dot done  ->left  ->next landed")
    for it in range(8, min(n_it - 1, 16)):
        print(f"{it:4d}  {int(t[2, it] - t[1, it]):6d} {int(t[3, it] - t[2, it]):6d} {int(t[11, it] - t[3, it]):6d} "
              f"{int(t[18, it] - t[11, it]):6d} {int(t[19, it] - t[18, it]):6d} {int(t[25, it] - t[19, it]):6d} {int(t[1, it + 1] - t[25, it]):6d}")
if t[29, 0] > 0:
    print(f"wall clock (ns): CTA 0 enter -> leave {int(t[29, 1] - t[29, 0])}; CTA 0 enter -> last CTA (block {int(t[29, 3])}) recorded the "
          f"scale {int(t[29, 2] - t[29, 0])}; CTA 0's cycle span start->end {int(t[31, 0] - t0)}")
if t[18, 0] > 0:
    print(f"set-up (cycles before 'start'): entry {int(t[18, 0] - t0)}, barriers initialised {int(t[18, 1] - t0)}, TMEM allocated {int(t[18, 2] - t0)}")
if t[19, 0] > 0:
    nb = int((t[19] > 0).sum())
    ent = (t[19, :nb] - t[29, 0]) / 1e3
    lea = (t[25, :nb] - t[29, 0]) / 1e3
    print(f"{nb} CTAs: entry (us after CTA 0's) min {ent.min():.2f} median {np.median(ent):.2f} max {ent.max():.2f}; roles done min {lea.min():.2f} "
          f"median {np.median(lea):.2f} max {lea.max():.2f}; busy time min {(lea - ent).min():.2f} median {np.median(lea - ent):.2f} max {(lea - ent).max():.2f}")
    order = np.argsort(lea)
    print("   slowest CTAs (block: entry, done):", [(int(b), round(float(ent[b]), 2), round(float(lea[b]), 2)) for b in order[-6:]])
    print("   fastest CTAs (block: entry, done):", [(int(b), round(float(ent[b]), 2), round(float(lea[b]), 2)) for b in order[:6]])
d = np.diff(t[9, :n_it])
if len(d) > 4:
    print("cycles between consecutive 'mma issued': median", int(np.median(d)), "mean", int(d.mean()))
