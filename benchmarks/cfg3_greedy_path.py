"""cfg 3 under ``optimize="auto"`` (SURVEY.md 8d asks for the explicit sweep AND the default optimiser).

Topology only, no data: builds the literal cfg-3 network's einsum string, runs the greedy optimiser that
``optimize="auto"`` selects for 1024 operands (contractn/einsum.py:213-220 hands the choice to opt_einsum; the planner
here mirrors its greedy rule) and prints the path's FLOPs and largest intermediate next to the explicit sweep's.
Result (this script, about 20 s): greedy 2.4e158 FLOPs, largest intermediate 2.7e157 elements -- the batch hub makes
every early pairing of two inputs look free, and the hub index then rides on every intermediate -- against 2.2e12 FLOPs
for the left-to-right sweep.  The default optimiser is therefore not a runnable configuration at literal size for any
implementation, the reference included; ``bench.py`` and the tests run cfg 3 on the sweep path (a tuple of tuples, the
one explicit form ``contract`` accepts).
"""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contractn_b200 import networks, planner  # noqa: E402


def main(sites=512, bond=512, phys=2, batch=4096):
    tn, sweep = networks.mps_tn(sites=sites, bond=2, phys=2, batch=2)  # topology; the shapes are scaled below
    einstr = tn.einsum_str
    shapes = [(phys, bond) if i in (0, sites - 1) else (phys, bond, bond) for i in range(sites)]
    shapes += [(batch, phys)] * sites
    dims = {}
    for lab, sh in zip(einstr.split("->")[0].split(","), shapes):
        dims.update(zip(lab, sh))
    out = {}
    for name, opt in (("auto", "auto"), ("sweep", sweep)):
        t0 = time.time()
        lst = planner.contract_path(einstr, *shapes, optimize=opt)[1]
        biggest = max(math.prod(dims[c] for c in step[2].split("->")[1]) for step in lst)
        out[name] = {"path_flops": float(planner.path_flops(lst, einstr, shapes)), "largest_intermediate_elements": float(biggest),
                     "path_search_s": round(time.time() - t0, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
