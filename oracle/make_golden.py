"""Generate tests/golden/*.npz by running the UNMODIFIED reference package.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

For every case it builds the network with the reference's own ``TN`` API (or
calls the reference's module-level ``contract``), records the einsum string,
the contraction list the reference executed, the operands and the reference's
outputs for both ``split_format`` settings.  ``opt_einsum`` is not installable
here, so ``oracle/oe_shim.py`` supplies it (path planning from
``contractn_b200.planner``; arithmetic is numpy's, driven by the reference's
own ``_core_contract`` / ``stabilize``).
"""
import json
import os
import random
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oe_shim  # noqa: E402

ref = oe_shim.import_reference()
import contractn.einsum as ref_einsum  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _encode_list(lst):
    return json.dumps(
        [[list(c[0]), sorted(c[1]), c[2], c[4] if c[4] else False] for c in lst], ensure_ascii=True
    )


def record(name, einstr, operands, optimize="auto", note="", **kw):
    shapes = tuple(tuple(o.shape) for o in operands)
    lst = ref_einsum._contract_path(einstr, shapes, optimize=optimize, memory_limit=kw.get("memory_limit"),
                                    use_blas=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            T, c = ref.contract(einstr, *operands, optimize=optimize, split_format=True, **kw)
            dense = ref.contract(einstr, *operands, optimize=optimize, split_format=False, **kw)
    payload = {
        "einstr": np.array(einstr),
        "optimize": np.array(json.dumps(optimize)),
        "memory_limit": np.array(json.dumps(kw.get("memory_limit"))),
        "contraction_list": np.array(_encode_list(lst)),
        "n_operands": np.array(len(operands)),
        "ref_T": np.asarray(T),
        "ref_c": np.asarray(c),
        "ref_dense": np.asarray(dense),
        "note": np.array(note),
    }
    # identical operands (README examples) are stored once
    uniq, index = [], []
    for o in operands:
        for j, u in enumerate(uniq):
            if u.shape == o.shape and u.dtype == o.dtype and np.array_equal(u, o):
                index.append(j)
                break
        else:
            index.append(len(uniq))
            uniq.append(o)
    payload["operand_index"] = np.array(index)
    for j, u in enumerate(uniq):
        payload[f"op{j}"] = u
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **payload)
    print(f"{name:28s} ops={len(operands):4d} steps={len(lst):4d} out={np.asarray(T).shape} c={float(c):.6g}")


def from_tn(name, tn, inputs=(), optimize="auto", note=""):
    einstr = tn.einsum_str
    packer = ref_einsum.make_arg_packer(tn)
    operands = packer(tn.params, inputs)
    # cross-check: the TN entry point gives the same result as the string entry
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            a = tn.contract(inputs=inputs, optimize=optimize, split_format=True)
            b = ref.contract(einstr, *operands, optimize=optimize, split_format=True)
    assert np.array_equal(a[0], b[0], equal_nan=True) and np.array_equal(a[1], b[1], equal_nan=True)
    record(name, einstr, operands, optimize=optimize, note=note)


def hub(n_vec, vec_fn, batch=None):
    tn = ref.TN()
    h = tn.add_copy_node(n_vec + 1)
    hb = tn.add_copy_node(n_vec + 1) if batch else None
    for i in range(n_vec):
        v = tn.add_dense_node(vec_fn(i))
        tn.connect_nodes(h, v, i, 0)
        if batch:
            tn.connect_nodes(hb, v, i, 1)
    return tn


def chain(n, vec, mat_fn):
    tn = ref.TN()
    prev = tn.add_dense_node(vec)
    for i in range(n):
        m = tn.add_dense_node(mat_fn(i))
        tn.connect_nodes(prev, m, -1, 0)
        prev = m
    return tn


def mps_ones(seed, max_bond, max_in, n):
    random.seed(seed)
    idims = [random.randint(1, max_in) for _ in range(n)]
    bdims = [random.randint(1, max_bond) for _ in range(n - 1)]
    tn = ref.TN()
    tn.add_dense_node(np.ones((idims[0], bdims[0])))
    for i in range(1, n - 1):
        tn.add_dense_node(np.ones((idims[i], bdims[i - 1], bdims[i])))
    tn.add_dense_node(np.ones((idims[-1], bdims[-1])))
    nodes = tn.nodes()
    for i in range(n - 1):
        tn.connect_nodes(nodes[i], nodes[i + 1], -1, -2 if i != n - 2 else -1)
    return tn


def batched_mps(rng, sites, bond, phys, batch, dtype):
    """cfg 3 in miniature: batch hub first, cores, inputs (batch, phys)."""
    tn = ref.TN()
    hubb = tn.add_copy_node(sites + 1)
    cores = []
    for i in range(sites):
        shape = (phys, bond) if i in (0, sites - 1) else (phys, bond, bond)
        cores.append(tn.add_dense_node(rng.uniform(0, 1, shape).astype(dtype)))
    for i in range(sites - 1):
        tn.connect_nodes(cores[i], cores[i + 1], -1, 1)
    for i in range(sites):
        x = tn.add_dense_node(rng.uniform(0.9, 1.1, (batch, phys)).astype(dtype))
        tn.connect_nodes(cores[i], x, 0, 1)
        tn.connect_nodes(hubb, x, i, 0)
    return tn


def peps(rng, L, D, dtype=np.float64, lo=0.0, hi=1.0):
    """cfg 5 in miniature: closed LxL grid, leg order (up, left, down, right)."""
    tn = ref.TN()
    grid = {}
    for r in range(L):
        for c in range(L):
            legs = [d for d, ok in (("u", r > 0), ("l", c > 0), ("d", r < L - 1), ("r", c < L - 1)) if ok]
            grid[r, c] = (tn.add_dense_node(rng.uniform(lo, hi, (D,) * len(legs)).astype(dtype)), legs)
    for r in range(L):
        for c in range(L):
            node, legs = grid[r, c]
            if c < L - 1:
                other, olegs = grid[r, c + 1]
                tn.connect_nodes(node, other, legs.index("r"), olegs.index("l"))
            if r < L - 1:
                other, olegs = grid[r + 1, c]
                tn.connect_nodes(node, other, legs.index("d"), olegs.index("u"))
    return tn


def cp_tn(rng, I, R, dtype=np.float64, signed=False):
    tn = ref.TN()
    h = tn.add_copy_node(3)
    for i in range(3):
        f = rng.normal(size=(I, R)) / np.sqrt(R) if signed else rng.uniform(0, 1, (I, R))
        m = tn.add_dense_node(f.astype(dtype))
        tn.connect_nodes(h, m, i, 1)
    return tn


def tucker_tn(rng, I, R, dtype=np.float64):
    tn = ref.TN()
    core = tn.add_dense_node(rng.uniform(0, 1, (R, R, R)).astype(dtype))
    for i in range(3):
        m = tn.add_dense_node(rng.uniform(0, 1, (I, R)).astype(dtype))
        tn.connect_nodes(core, m, i, 1)
    return tn


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20211214)

    # README / notebook known answers (README.md:33,74-76)
    from_tn("readme_hub", hub(100, lambda i: np.array([1, 0.99])), note="README.md:18-34")
    from_tn("readme_chain", chain(1000, np.ones((3,)), lambda i: np.ones((3, 3))), note="README.md:62-77")
    cp, tk = ref.TN(), ref.TN()
    ch, th = cp.add_copy_node(3), tk.add_dense_node(np.ones((4, 4, 4)))
    for i in range(3):
        a, b = cp.add_dense_node(np.eye(4, 10)), tk.add_dense_node(np.eye(4, 10))
        cp.connect_nodes(ch, a, i, 0)
        tk.connect_nodes(th, b, i, 0)
    from_tn("readme_cp", cp, note="README.md:40-58")
    from_tn("readme_tucker", tk, note="README.md:40-58")

    # the reference's only numeric test (tests/test_einsum.py:18-64), four draws
    for k, (seed, mb, mi, n) in enumerate([(1, 10, 6, 2), (2, 7, 4, 4), (3, 10, 6, 6), (5, 3, 1, 5)]):
        from_tn(f"mps_ones_{k}", mps_ones(seed, mb, mi, n), note="tests/test_einsum.py:18-64")

    # cfg 1 with random entries
    from_tn("chain_random_200", chain(200, rng.uniform(0.5, 1.5, (3,)), lambda i: rng.uniform(0.5, 1.5, (3, 3))))
    from_tn("chain_signed_64", chain(64, rng.normal(size=(5,)), lambda i: rng.normal(size=(5, 5))))
    # cfg 2 in miniature (batched hub), f64 and f32
    from_tn("hub_batched_f64", hub(24, lambda i: rng.uniform(0.9, 1.1, (2, 96)), batch=True))
    from_tn("hub_batched_f32", hub(24, lambda i: rng.uniform(0.9, 1.1, (2, 96)).astype(np.float32), batch=True))
    from_tn("hub_small_values", hub(40, lambda i: rng.uniform(1e-4, 2e-4, (3, 50)), batch=True),
            note="every operand ~1e-4: the raw product underflows the 1e-7 threshold, the rescaled one does not")
    from_tn("hub_large_values", hub(60, lambda i: rng.uniform(1e6, 2e6, (2, 40)), batch=True))
    from_tn("hub_three", hub(3, lambda i: rng.uniform(0.1, 1.0, (7, 33)), batch=True))
    # cfg 3 in miniature
    from_tn("mps_batched_f32", batched_mps(rng, 8, 6, 2, 16, np.float32))
    from_tn("mps_batched_f64", batched_mps(rng, 8, 6, 2, 16, np.float64))
    from_tn("mps_batched_f64_wide", batched_mps(rng, 5, 40, 2, 130, np.float64))
    # cfg 4 in miniature
    from_tn("cp_f64", cp_tn(rng, 24, 16))
    from_tn("cp_signed_f64", cp_tn(rng, 20, 12, signed=True))
    from_tn("tucker_f64", tucker_tn(rng, 18, 8))
    from_tn("tucker_f32", tucker_tn(rng, 18, 8, np.float32))
    # cfg 5 in miniature
    from_tn("peps_3x3_d2", peps(rng, 3, 2))
    from_tn("peps_4x4_d3", peps(rng, 4, 3))
    from_tn("peps_4x4_d2_f32", peps(rng, 4, 2, np.float32))
    from_tn("peps_5x5_d2_signed", peps(rng, 5, 2, lo=-1.0, hi=1.0))

    # module-level contract(): edge cases of the executor
    z = rng.uniform(0, 1, (4, 5))
    record("zero_operand", "ab,bc,cd->ad", [z, np.zeros((5, 6)), rng.uniform(0, 1, (6, 3))],
           note="zero tensor: norm 0 <= 1e-7 -> no rescale (einsum.py:102)")
    record("tiny_norm_skip", "ab,bc,cd->ad",
           [1e-6 * rng.uniform(0, 1, (3, 4)), 1e-6 * rng.uniform(0, 1, (4, 5)), rng.uniform(0, 1, (5, 2))],
           note="first product has sum|T| <= 1e-7 and is left unscaled")
    record("single_operand", "ab->ba", [rng.uniform(0, 1, (4, 6))], note="1-tensor network is stabilised too")
    record("single_trace", "aab->b", [rng.uniform(0, 1, (5, 5, 3))])
    record("diag_repeat", "aab,bc,ca->a", [rng.uniform(0, 1, (4, 4, 5)), rng.uniform(0, 1, (5, 6)),
                                           rng.uniform(0, 1, (6, 4))])
    record("scalar_out", "ab,ab->", [rng.uniform(0, 1, (7, 9)), rng.uniform(0, 1, (7, 9))])
    record("scalar_negative", "a,a->", [np.array([1.0, -3.0, 0.5]), np.array([1.0, 2.0, 1.0])],
           note="sign lives in the mantissa")
    record("outer", "a,b,c->abc", [rng.uniform(0, 1, (3,)), rng.uniform(0, 1, (4,)), rng.uniform(0, 1, (5,))])
    record("implicit_sum", "ab,cd->ad", [rng.uniform(0, 1, (3, 4)), rng.uniform(0, 1, (5, 6))],
           note="b and c are summed out of one operand each")
    record("broadcast_dim", "ab,bc->ac", [rng.uniform(0, 1, (3, 1)), rng.uniform(0, 1, (4, 5))])
    record("batch_matmul", "bij,bjk->bik", [rng.uniform(0, 1, (6, 5, 7)), rng.uniform(0, 1, (6, 7, 4))])
    record("transposed_out", "abc,cd->dba", [rng.uniform(0, 1, (3, 4, 5)), rng.uniform(0, 1, (5, 6))])
    record("explicit_path", "ab,bc,cd,de->ae",
           [rng.uniform(0, 1, (3, 4)), rng.uniform(0, 1, (4, 5)), rng.uniform(0, 1, (5, 6)),
            rng.uniform(0, 1, (6, 2))], optimize=((2, 3), (0, 1), (0, 1)))
    record("nary_memory_limit", "ab,bc,cd,de->ae",
           [rng.uniform(0, 1, (3, 4)), rng.uniform(0, 1, (4, 5)), rng.uniform(0, 1, (5, 6)),
            rng.uniform(0, 1, (6, 2))], optimize="optimal", memory_limit=5,
           note="memory limit forces one n-ary einsum step")
    record("mixed_dtype", "ab,bc->ac", [rng.uniform(0, 1, (3, 4)).astype(np.float32), rng.uniform(0, 1, (4, 5))])
    record("int_operands", "ab,bc->ac", [np.arange(12).reshape(3, 4), np.arange(20).reshape(4, 5)])
    record("strided_views", "ab,bc->ac", [rng.uniform(0, 1, (8, 6))[::2, ::3].T.copy().T,
                                          np.asfortranarray(rng.uniform(0, 1, (2, 5)))])
    record("big_k_f64", "ab,bc->ac", [rng.uniform(0, 1, (33, 700)), rng.uniform(0, 1, (700, 17))])
    record("big_k_f32", "ab,bc->ac", [rng.uniform(0, 1, (33, 700)).astype(np.float32),
                                      rng.uniform(0, 1, (700, 17)).astype(np.float32)])
    print("done ->", OUT)


if __name__ == "__main__":
    main()
