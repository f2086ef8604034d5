"""Record the einsum strings the UNMODIFIED reference TN compiler produces for a
set of construction recipes -> tests/golden/tn_strings.json.  TEST INFRASTRUCTURE;
build container only (needs /root/reference):  python -m oracle.make_tn_golden
"""
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oe_shim  # noqa: E402

ref = oe_shim.import_reference()
from oracle.make_tn_golden_replay import replay  # noqa: E402


def random_recipe(rng):
    rec, legs = [], []
    n = rng.randint(2, 10)
    for i in range(n):
        r = rng.random()
        if r < 0.25 and i < n - 1:
            deg = rng.randint(2, 5)
            rec.append(["copy", deg])
            legs += [(i, l, None) for l in range(deg)]
        else:
            nd = rng.randint(1, 4)
            shp = [rng.choice([2, 3]) for _ in range(nd)]
            rec.append(["input" if r > 0.9 else "dense", shp])
            legs += [(i, l, shp[l]) for l in range(nd)]
    rng.shuffle(legs)
    hubdim, used, conns = {}, set(), []
    for a in range(len(legs)):
        for b in range(a + 1, len(legs)):
            if a in used or b in used:
                continue
            (n1, l1, d1), (n2, l2, d2) = legs[a], legs[b]
            if n1 == n2 or (d1 is None and d2 is None):
                continue
            # copy nodes precede the dense nodes they attach to (SURVEY finding 0.7)
            if (d1 is None and n1 > n2) or (d2 is None and n2 > n1):
                continue
            if d1 is None:
                if hubdim.get(n1, d2) != d2:
                    continue
                hubdim[n1] = d2
            elif d2 is None:
                if hubdim.get(n2, d1) != d1:
                    continue
                hubdim[n2] = d1
            elif d1 != d2:
                continue
            if rng.random() < 0.5:
                used |= {a, b}
                conns.append(["conn", n1, n2, l1, l2])
    rec += conns
    # cutting a bond of a copy node makes the reference's relabelling depend on
    # set iteration order (a tie in Counter.most_common over a set of node
    # names, ctn.py:290-295), i.e. on PYTHONHASHSEED: only dense-dense bonds here
    plain = [c for c in conns if rec[c[1]][0] != "copy" and rec[c[2]][0] != "copy"]
    if plain and rng.random() < 0.3:
        c = rng.choice(plain)
        rec.append(["rm", c[1], c[3]])
    return rec


def fixed_recipes():
    out = {}
    out["readme_cp"] = [["copy", 3]] + [x for i in range(3) for x in (["dense", [4, 10]], ["conn", 0, i + 1, i, 0])]
    out["readme_tucker"] = [["dense", [4, 4, 4]]] + [x for i in range(3) for x in (["dense", [4, 10]], ["conn", 0, i + 1, i, 0])]
    chain = [["dense", [3]]]
    for i in range(80):
        chain += [["dense", [3, 3]], ["conn", i, i + 1, -1, 0]]
    out["chain_80"] = chain
    hub = [["copy", 61]]
    for i in range(60):
        hub += [["dense", [2]], ["conn", 0, i + 1, i, 0]]
    out["hub_60"] = hub
    bh = [["copy", 31], ["copy", 31]]
    for i in range(30):
        bh += [["dense", [2, 8]], ["conn", 0, i + 2, i, 0], ["conn", 1, i + 2, i, 1]]
    out["hub_batched_30"] = bh
    return out


def main():
    rng = random.Random(20211214)
    cases = []
    for name, rec in fixed_recipes().items():
        cases.append({"name": name, "recipe": rec, "einsum_str": replay(ref.TN, rec).einsum_str})
    k = 0
    while k < 150:
        rec = random_recipe(rng)
        try:
            s = replay(ref.TN, rec).einsum_str
        except Exception:
            continue
        cases.append({"name": f"random_{k}", "recipe": rec, "einsum_str": s})
        k += 1
    path = os.path.join(ROOT, "tests", "golden", "tn_strings.json")
    with open(path, "w") as f:
        json.dump(cases, f, ensure_ascii=True)
    print(len(cases), "cases ->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
