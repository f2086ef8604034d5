"""Parity of the CUDA path against the oracle / the reference's golden vectors.

All tests here need a B200 (``-m gpu``) and call through the C ABI
(``ctn_plan_execute_host`` / ``ctn_plan_execute``).  Tolerances are north_star's:
float64 values rtol 1e-12 and log-scale atol 1e-9; float32 values rtol 1e-5.
"""
import numpy as np
import pytest

from oracle import contract_oracle as oracle
from contractn_b200 import planner

from _cases import Golden, golden_names, assert_split_close, tolerances

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names())
def test_golden_split_and_dense(gpu_engine, name):
    g = Golden(name)
    T, c = gpu_engine.core_contract(g.operands, g.contraction_list)
    assert T.dtype == g.ref_T.dtype
    assert np.asarray(c).dtype == np.float64 and np.asarray(c).shape == ()
    assert_split_close(T, c, g.ref_T, g.ref_c, g.is_f32, g.signed, what=name)
    # user-level entry: same string, same optimize -> same path -> same values
    kw = {} if g.memory_limit is None else {"memory_limit": g.memory_limit}
    with np.errstate(over="ignore"):
        dense = gpu_engine.contract(g.einstr, *g.operands, optimize=g.optimize, **kw)
    ref = np.asarray(g.ref_dense, dtype=np.float64)
    got = np.asarray(dense, dtype=np.float64)
    assert got.shape == ref.shape
    if np.all(np.isfinite(ref)):
        if g.signed:
            assert np.linalg.norm((got - ref).ravel()) <= tolerances(g.is_f32)["rtol"] * 10 * np.linalg.norm(ref.ravel())
        else:
            np.testing.assert_allclose(got, ref, rtol=1e-4 if g.is_f32 else 1e-9)
    else:
        assert np.array_equal(np.isinf(got), np.isinf(ref))  # README: [inf inf inf]


@pytest.mark.parametrize("name", golden_names())
def test_golden_exact_step_mode(gpu_engine, name):
    """CTN_PLAN_EXACT_STEPS: no fusion, one kernel per pairwise step."""
    g = Golden(name)
    dtype = np.float32 if g.is_f32 else np.float64
    ops = [np.ascontiguousarray(np.asarray(o, dtype=dtype)) for o in g.operands]
    flags = gpu_engine.CTN_PLAN_EXACT_STEPS | gpu_engine.CTN_PLAN_NO_INTERPRETER
    plan = gpu_engine.Plan(g.contraction_list, [o.shape for o in ops], dtype, flags=flags)
    assert "interp" not in plan.launch_kinds
    T, c, fl = plan.execute_host(ops)
    assert fl == 0
    assert_split_close(T, c, g.ref_T, g.ref_c, g.is_f32, g.signed, what=name)


@pytest.mark.parametrize("name", ["hub_batched_f64", "mps_batched_f32", "cp_f64", "peps_4x4_d3", "readme_chain"])
def test_torch_device_resident_path(gpu_engine, name):
    import torch

    g = Golden(name)
    ops = [torch.as_tensor(np.ascontiguousarray(o)).cuda() for o in g.operands]
    T, c = gpu_engine.core_contract(ops, g.contraction_list)
    assert T.is_cuda and c.is_cuda and c.dim() == 0
    assert_split_close(T.cpu().numpy(), c.item(), g.ref_T, g.ref_c, g.is_f32, g.signed, what=name)
    # and the operands were not touched
    for t, o in zip(ops, g.operands):
        assert np.array_equal(t.cpu().numpy(), o)
    # split_format=False: destabilize (einsum.py:110-114) runs in the plan's last kernel (CTN_OUT_DENSE)
    kw = {} if g.memory_limit is None else {"memory_limit": g.memory_limit}
    dense = gpu_engine.contract(g.einstr, *ops, optimize=g.optimize, **kw)
    assert dense.is_cuda and dense.dtype == ops[0].dtype
    ref = np.asarray(g.ref_dense, dtype=np.float64)
    got = dense.cpu().numpy().astype(np.float64)
    if np.all(np.isfinite(ref)):
        np.testing.assert_allclose(got, ref, rtol=1e-4 if g.is_f32 else 1e-9)
    else:
        assert np.array_equal(np.isinf(got), np.isinf(ref))


def test_readme_examples_through_tn_api(gpu_engine):
    """README.md:18-34 and :62-77 through contractn_b200.TN."""
    from contractn_b200 import TN

    tn = TN()
    hub = tn.add_copy_node(101)
    for i in range(100):
        tn.connect_nodes(hub, tn.add_dense_node(np.array([1, 0.99])), i, 0)
    np.testing.assert_allclose(tn.contract(), [1.0, 0.36603234], rtol=1e-8)
    T, c = tn.contract(split_format=True)
    np.testing.assert_allclose(T, [1.46409418, 0.53590582], rtol=1e-8)
    assert abs(float(c) - (-0.38123674)) < 1e-8

    tn = TN()
    prev = tn.add_dense_node(np.ones((3,)))
    for i in range(1000):
        mat = tn.add_dense_node(np.ones((3, 3)))
        tn.connect_nodes(prev, mat, -1, 0)
        prev = mat
    with np.errstate(over="ignore"):
        assert np.all(np.isinf(tn.contract()))
    T, c = tn.contract(split_format=True)
    np.testing.assert_allclose(T, [1.0, 1.0, 1.0], rtol=1e-12)
    assert abs(float(c) - 1098.61228867) < 1e-8
    assert abs(float(c) - 1000 * np.log(3.0)) < 1e-9


def test_reference_mps_test_on_gpu(gpu_engine):
    """tests/test_einsum.py:18-64 re-run against the engine: all-ones MPS,
    log of every entry equals the sum of log bond dimensions."""
    import math
    import random

    from contractn_b200 import TN

    for seed in range(1, 6):
        for n in range(2, 7):
            random.seed(seed * 100 + n)
            idims = [random.randint(1, 6) for _ in range(n)]
            bdims = [random.randint(1, 10) for _ in range(n - 1)]
            tn = TN()
            tn.add_dense_node(np.ones((idims[0], bdims[0])))
            for i in range(1, n - 1):
                tn.add_dense_node(np.ones((idims[i], bdims[i - 1], bdims[i])))
            tn.add_dense_node(np.ones((idims[-1], bdims[-1])))
            nodes = tn.nodes()
            for i in range(n - 1):
                tn.connect_nodes(nodes[i], nodes[i + 1], -1, -2 if i != n - 2 else -1)
            want = sum(math.log(b) for b in bdims)
            for split in (False, True):
                out = tn.contract(split_format=split)
                logv = np.log(out[0]) + out[1] if split else np.log(out)
                assert logv.shape == tuple(idims)
                np.testing.assert_allclose(logv, want, rtol=1e-12, atol=1e-12)


def test_fused_chain_matches_stepwise_and_oracle(gpu_engine):
    """cfg 2 shape, medium size: n-ary fused kernel == per-step kernels == oracle."""
    rng = np.random.default_rng(1002)
    n, B = 100, 1 << 14
    ops = [rng.uniform(0.9, 1.1, (2, B)) for _ in range(n)]
    einstr = ",".join(["ab"] * n) + "->ab"
    lst = planner.contract_path(einstr, *[o.shape for o in ops])[1]
    T0, c0 = oracle.core_contract(ops, lst)
    fused = gpu_engine.Plan(lst, [o.shape for o in ops], np.float64)
    assert fused.launch_kinds == ("chain", "materialize")
    T1, c1, fl = fused.execute_host(ops)
    assert fl == 0
    assert_split_close(T1, c1, T0, c0, False, what="fused")
    step = gpu_engine.Plan(lst, [o.shape for o in ops], np.float64, flags=gpu_engine.CTN_PLAN_EXACT_STEPS)
    T2, c2, _ = step.execute_host(ops)
    assert_split_close(T2, c2, T0, c0, False, what="stepwise")
    # bitwise reproducible from run to run (deterministic reductions)
    T3, c3, _ = fused.execute_host(ops)
    assert np.array_equal(T1, T3) and c1 == c3


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fused_chain_extreme_ranges(gpu_engine, dtype):
    """Products that under/overflow the dtype without the per-element exponent.
    (Operands must keep sum|T_hat * v| above the reference's 1e-7 threshold,
    einsum.py:102, or the reference itself stops rescaling and underflows.)"""
    rng = np.random.default_rng(7)
    for lo, hi, n in [(1e-9, 2e-9, 64), (1e25, 3e25, 48), (1e-3, 1e3, 300)]:
        ops = [rng.uniform(lo, hi, (3, 1000)).astype(dtype) for _ in range(n)]
        einstr = ",".join(["ab"] * n) + "->ab"
        lst = planner.contract_path(einstr, *[o.shape for o in ops])[1]
        T, c = gpu_engine.core_contract(ops, lst)
        # truth from log-space arithmetic in float64
        logs = sum(np.log(np.asarray(o, dtype=np.float64)) for o in ops)
        m = logs.max()
        mean = np.exp(logs - m).mean()
        want_c = m + np.log(mean)
        want_T = np.exp(logs - m) / mean
        f32 = dtype == np.float32
        assert abs(float(c) - want_c) <= (2e-3 if f32 else 1e-8)
        big = want_T > 1e-30
        np.testing.assert_allclose(np.asarray(T, np.float64)[big], want_T[big], rtol=2e-4 if f32 else 1e-9)


def test_fused_chain_degenerate_falls_back_to_exact_steps(gpu_engine):
    """A zero operand inside a hub: the reference keeps the log-scale it had
    accumulated before the product vanished; the fused kernel flags the case
    and the engine re-runs it step by step (still on the GPU)."""
    rng = np.random.default_rng(3)
    ops = [rng.uniform(0.5, 1.5, (2, 4096)) for _ in range(10)]
    ops[6] = np.zeros((2, 4096))
    einstr = ",".join(["ab"] * 10) + "->ab"
    lst = planner.contract_path(einstr, *[o.shape for o in ops])[1]
    T0, c0 = oracle.core_contract(ops, lst)
    fused = gpu_engine.Plan(lst, [o.shape for o in ops], np.float64)
    _, _, fl = fused.execute_host(ops)
    assert fl & gpu_engine.CTN_RESULT_DEGENERATE
    T, c = gpu_engine.core_contract(ops, lst)
    assert np.array_equal(T, T0)
    assert abs(float(c) - float(c0)) < 1e-9


def test_cfg2_full_size_properties(gpu_engine):
    """BASELINE config 2 at full size (100 x (2, 2^20) float64, 1.7 GB):
    README known answer broadcast over the batch, and the canonical-split
    property against a float64 numpy product."""
    B = 1 << 20
    n = 100
    einstr = ",".join(["ab"] * n) + "->ab"
    lst = planner.contract_path(einstr, *[(2, B)] * n)[1]
    plan = gpu_engine.Plan(lst, [(2, B)] * n, np.float64)
    assert plan.launch_kinds == ("chain", "materialize")
    kat = np.ascontiguousarray(np.broadcast_to(np.array([[1.0], [0.99]]), (2, B)))
    T, c, fl = plan.execute_host([kat] * n)
    assert fl == 0
    want = np.array([1.0, 0.99 ** 100])
    np.testing.assert_allclose(T[:, 0] * np.exp(c), want, rtol=1e-12)
    assert np.all(T[0] == T[0, 0]) and np.all(T[1] == T[1, 0])
    rng = np.random.default_rng(1002)
    ops = [rng.uniform(0.9, 1.1, (2, B)) for _ in range(n)]
    T, c, fl = plan.execute_host(ops)
    prod = np.ones((2, B))
    for o in ops:
        prod *= o
    wT, wc = oracle.canonical_split(prod)
    np.testing.assert_allclose(T, wT, rtol=1e-12)
    assert abs(c - wc) < 1e-9
    assert abs(np.abs(T).mean() - 1.0) < 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_batched_mps_under_the_default_optimizer(gpu_engine, dtype):
    """cfg 3's network at a size where ``optimize="auto"`` is runnable (at literal size its greedy path needs
    1e157 elements, benchmarks/cfg3_greedy_path.py): default optimiser and explicit sweep give the oracle's
    result, each on its own path."""
    from contractn_b200 import networks

    tn, sweep = networks.mps_tn(sites=6, bond=8, phys=2, batch=16, dtype=dtype, seed=9)
    ops = [np.asarray(p) for p in tn.params]
    shapes = tuple(o.shape for o in ops)
    for opt in ("auto", sweep):
        lst = planner.contract_path(tn.einsum_str, *shapes, optimize=opt)[1]
        T, c = gpu_engine.core_contract(ops, lst)
        rT, rc = oracle.core_contract(ops, lst)
        assert_split_close(T, c, rT, rc, dtype == np.float32, False, what="mps/%s" % (opt if opt == "auto" else "sweep"))


@pytest.mark.parametrize("einstr,shapes", [
    ("ab,bc,cd->ad", [(30, 40), (40, 0), (0, 50)]),
    ("ab,bc,cd->ad", [(30, 40), (40, 50), (50, 0)]),
    ("ab,bc,cd,de->ae", [(30, 40), (40, 50), (50, 0), (0, 20)]),
])
def test_zero_extent_operands(gpu_engine, einstr, shapes):
    """einsum.py:97-106 with empty tensors: zeros (or an empty result) and the log-scale of the steps among the
    non-empty operands, numpy and torch, split and dense."""
    import torch

    rng = np.random.default_rng(4)
    ops = [rng.uniform(0.5, 1.5, s) for s in shapes]
    lst = planner.contract_path(einstr, *shapes)[1]
    with np.errstate(invalid="ignore"):
        want_T, want_c = oracle.core_contract(ops, lst)
    T, c = gpu_engine.core_contract(ops, lst)
    assert T.shape == want_T.shape and T.dtype == want_T.dtype and not np.any(T)
    assert abs(float(c) - float(want_c)) < 1e-9
    dense = gpu_engine.contract(einstr, *ops)
    assert dense.shape == want_T.shape and not np.any(dense)
    tT, tc = gpu_engine.core_contract([torch.as_tensor(o).cuda() for o in ops], lst)
    assert tuple(tT.shape) == want_T.shape and tT.is_cuda and not bool(tT.any())
    assert abs(float(tc) - float(want_c)) < 1e-9
