"""Wall time of the README examples through the public API (what the reference's only published
timing measures: `%time tn.contract()` = 6.85 ms for the order-101 hub, README.md:34)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contractn_b200 import TN

def hub():
    tn = TN()
    c = tn.add_copy_node(101)
    for i in range(100):
        tn.connect_nodes(c, tn.add_dense_node(np.array([1, 0.99])), i, 0)
    return tn

def chain():
    tn = TN()
    prev = tn.add_dense_node(np.ones((3,)))
    for i in range(1000):
        m = tn.add_dense_node(np.ones((3, 3)))
        tn.connect_nodes(prev, m, -1, 0)
        prev = m
    return tn

for name, build, kw in (("readme_hub", hub, {}), ("readme_chain", chain, {"split_format": True})):
    tn = build()
    t0 = time.perf_counter(); out = tn.contract(**kw); first = time.perf_counter() - t0
    times = []
    for _ in range(20):
        t0 = time.perf_counter(); out = tn.contract(**kw); times.append(time.perf_counter() - t0)
    fun = tn.make_contract_fun(**kw); params = tn.params
    t2 = []
    for _ in range(20):
        t0 = time.perf_counter(); fun(params, ()); t2.append(time.perf_counter() - t0)
    print(json.dumps({"example": name, "first_call_ms": first * 1e3, "tn_contract_ms_median": float(np.median(times)) * 1e3,
                      "contract_fun_ms_median": float(np.median(t2)) * 1e3, "result": np.asarray(out[0] if isinstance(out, tuple) else out).tolist()[:3],
                      "reference_published_ms": 6.85 if name == "readme_hub" else None}), flush=True)
