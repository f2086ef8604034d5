import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from contractn_b200 import engine, planner
shapes = [(1024, 256), (1024, 256)]
ops = [torch.rand(s, dtype=torch.float64, device="cuda") for s in shapes]
lst = planner.contract_path("ca,ba->cab", *shapes)[1]
plan = engine.Plan(lst, shapes, np.float64)
for _ in range(3): plan.execute_device(ops)
torch.cuda.synchronize()
