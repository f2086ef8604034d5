"""Debug: per-step clock stamps of the interpreter kernel (library built with -DCTN_INTERP_PROFILE as
contractn_b200/libctn_prof.so; see the PROF() marks in csrc/kernels.cu)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contractn_b200 import engine
engine._LIB_PATH = os.path.join(os.path.dirname(engine._LIB_PATH), "libctn_prof.so")
from contractn_b200 import planner
from benchmarks.configs import CONFIGS
tn, optimize = CONFIGS["cfg1"][1]()
params = tn.params
shapes = tuple(tuple(p.shape) for p in params)
lst = planner.cached_contract_list(tn.einsum_str, shapes, optimize=optimize)
plan = engine.Plan(lst, shapes, params[0].dtype)
dev = [torch.as_tensor(p).cuda() for p in params]
for _ in range(3):
    plan.execute_device(dev)
torch.cuda.synchronize()
n = len(lst)
buf = np.zeros((n, 8), dtype=np.int64)
lib = engine.load_library()
lib.ctn_debug_interp_profile.argtypes = [ctypes.c_void_p, ctypes.c_int]
print("rc", lib.ctn_debug_interp_profile(buf.ctypes.data, n))
t = buf - buf[0, 0]
print("first rows\n", t[:12])
print("rows 500..506\n", t[500:506])
d = np.diff(t[:, 3])
print("publish-to-publish cycles: median", np.median(d), "mean", d.mean())
print("wait_done - prev publish: median", np.median(t[1:, 1] - t[:-1, 3]))
print("compute (wait_done -> before warp_sum): median", np.median(t[:, 2] - t[:, 1]))
print("reduce+div+publish: median", np.median(t[:, 3] - t[:, 2]))
print("fence+alpha (1->4): median", np.median(t[:, 4] - t[:, 1]))
print("element loop (4->2): median", np.median(t[:, 2] - t[:, 4]))
print("warp_sum (2->5): median", np.median(t[:, 5] - t[:, 2]))
print("div+publish (5->3): median", np.median(t[:, 3] - t[:, 5]))
print("total cycles", t[-1, 3])
