"""Interpreter kernel (runs of tiny steps in one launch) against the oracle: the short pairwise path
(<= 32 outputs, <= 1 summed mode, indices decoded before the wait), the general pairwise path, n-ary
steps, operands staged in shared memory (strided views, shared operands), chains longer than one
launch.  Reference semantics: contractn/einsum.py:326-393 (loop) and :89-107 (stabilize)."""
import numpy as np
import pytest

from oracle import contract_oracle as oracle
from contractn_b200 import planner

from _cases import assert_split_close

pytestmark = pytest.mark.gpu


def _chain(n, d, rng, dtype):
    syms = [planner.get_symbol(i) for i in range(n + 1)]
    einstr = syms[0] + "," + ",".join(syms[i] + syms[i + 1] for i in range(n)) + "->" + syms[n]
    ops = [rng.uniform(0.5, 1.5, (d,)).astype(dtype)] + [rng.uniform(0.5, 1.5, (d, d)).astype(dtype) for _ in range(n)]
    return einstr, ops


def _check(gpu_engine, einstr, ops, optimize="auto", f32=False, only_interp=True):
    lst = planner.contract_path(einstr, *[o.shape for o in ops], optimize=optimize)[1]
    ref_T, ref_c = oracle.core_contract(ops, lst)
    T, c = gpu_engine.core_contract(ops, lst)
    plan = gpu_engine.get_plan(lst, tuple(o.shape for o in ops), ops[0].dtype,
                               strides=tuple(gpu_engine._elem_strides(o) for o in ops))
    if only_interp:
        assert set(plan.launch_kinds) == {"interp"}, plan.launch_kinds
    assert_split_close(T, c, ref_T, ref_c, f32, what=einstr[:40])
    return plan


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,d", [(50, 3), (200, 2), (64, 5), (30, 7)])
def test_dependent_chain(gpu_engine, dtype, n, d):
    einstr, ops = _chain(n, d, np.random.default_rng(n * 10 + d), dtype)
    _check(gpu_engine, einstr, ops, f32=dtype == np.float32)


def test_chain_longer_than_one_launch(gpu_engine):
    """> 4096 steps: the second launch consumes an intermediate (and its pending scale) of the first."""
    einstr, ops = _chain(4500, 3, np.random.default_rng(3), np.float64)
    plan = _check(gpu_engine, einstr, ops)
    assert plan.launch_kinds.count("interp") >= 2


@pytest.mark.parametrize("shape_a,shape_b,einstr", [
    ((4, 8), (8, 8), "ab,bc->ac"),        # 32 outputs: the short path's upper edge
    ((3, 11), (11, 11), "ab,bc->ac"),     # 33 outputs: general pairwise path
    ((2, 3, 4), (3, 4, 5), "abc,bcd->ad"),  # two summed modes
    ((5, 6), (5, 6), "ab,ab->"),          # full reduction, scalar output
    ((4, 5), (5,), "ab,b->ab"),           # no summed mode, broadcast
    ((6,), (7,), "a,b->ab"),              # outer product
    ((3, 4, 5), (5, 4), "abc,cb->a"),     # transposed contraction
])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_single_tiny_steps(gpu_engine, shape_a, shape_b, einstr, dtype):
    rng = np.random.default_rng(len(einstr))
    ops = [rng.uniform(-1, 1, shape_a).astype(dtype), rng.uniform(-1, 1, shape_b).astype(dtype)]
    lst = planner.contract_path(einstr, *[o.shape for o in ops])[1]
    ref_T, ref_c = oracle.core_contract(ops, lst)
    T, c = gpu_engine.core_contract(ops, lst)
    assert_split_close(T, c, ref_T, ref_c, dtype == np.float32, signed=True, what=einstr)


def test_strided_views_and_shared_operands(gpu_engine):
    """Operands staged into shared memory keep their strides (the whole address span is copied);
    one buffer bound to several slots is staged once."""
    import torch

    rng = np.random.default_rng(11)
    big = torch.as_tensor(rng.uniform(0.5, 1.5, (6, 9))).cuda()
    view = big[::2, ::3]                       # (3, 3), strides (18, 3)
    tview = torch.as_tensor(rng.uniform(0.5, 1.5, (3, 3))).cuda().t()  # transposed view
    v = torch.as_tensor(rng.uniform(0.5, 1.5, (3,))).cuda()
    ops = [v, view, tview, view, tview, view]
    einstr = "a,ab,bc,cd,de,ef->f"
    lst = planner.contract_path(einstr, *[tuple(o.shape) for o in ops])[1]
    ref_T, ref_c = oracle.core_contract([o.cpu().numpy() for o in ops], lst)
    T, c = gpu_engine.core_contract(ops, lst)
    assert_split_close(T.cpu().numpy(), float(c), ref_T, ref_c, False, what="strided")
    assert torch.equal(big[::2, ::3], view)   # operands are never written


def test_nary_tiny_steps(gpu_engine):
    rng = np.random.default_rng(5)
    ops = [rng.uniform(0.5, 1.5, (4, 5)) for _ in range(3)] + [rng.uniform(0.5, 1.5, (5, 3))]
    einstr = "ab,ab,ab,bc->ac"
    _check(gpu_engine, einstr, ops, only_interp=False)


def test_tiny_norms_skip_rescale_inside_interpreter(gpu_engine):
    """sum|T| <= 1e-7 leaves tensor and log-scale untouched (contractn/einsum.py:102)."""
    rng = np.random.default_rng(8)
    einstr, ops = _chain(6, 3, rng, np.float64)
    ops = [o * 1e-5 for o in ops]  # every product has sum|T| far below the threshold
    lst = planner.contract_path(einstr, *[o.shape for o in ops])[1]
    ref_T, ref_c = oracle.core_contract(ops, lst)
    T, c = gpu_engine.core_contract(ops, lst)
    assert float(ref_c) == 0.0 and float(c) == 0.0
    np.testing.assert_allclose(T, ref_T, rtol=1e-12)


def test_repeated_calls_reuse_host_bundle_but_read_fresh_values(gpu_engine):
    """The host path caches per-operand bookkeeping between calls on the same array objects; values
    are always read from the caller's buffers, and a replaced operand drops the cache."""
    rng = np.random.default_rng(21)
    einstr, ops = _chain(40, 3, rng, np.float64)
    lst = planner.contract_path(einstr, *[o.shape for o in ops])[1]
    for trial in range(3):
        ref_T, ref_c = oracle.core_contract(ops, lst)
        T, c = gpu_engine.core_contract(ops, lst)
        assert_split_close(T, c, ref_T, ref_c, False, what=f"trial {trial}")
        ops[7][...] = rng.uniform(0.5, 1.5, ops[7].shape)      # in-place update, same object
    ops[3] = np.asfortranarray(rng.uniform(0.5, 1.5, (3, 3)))  # new object, different strides
    ref_T, ref_c = oracle.core_contract(ops, lst)
    T, c = gpu_engine.core_contract(ops, lst)
    assert_split_close(T, c, ref_T, ref_c, False, what="replaced operand")


DOT_CASES = [
    ("ab,ab->", [(300, 400), (300, 400)]),           # inner product, one fused summed mode
    ("ab,ba->", [(300, 400), (400, 300)]),           # two summed modes that do not fuse
    ("abk,bk->a", [(4, 30, 2000), (30, 2000)]),      # several outputs
    ("ak,ak->a", [(8, 50000), (8, 50000)]),          # batch index kept, long sum
    ("k,k,k->", [(100000,), (100000,), (100000,)]),  # n-ary
    ("abc->", [(40, 50, 60)]),                       # single-operand full reduction
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("einstr,shapes", DOT_CASES, ids=[c[0] for c in DOT_CASES])
def test_dot_class_steps(gpu_engine, einstr, shapes, dtype):
    """Few outputs, each a long sum (DOT class of einsum.py:347-384): the grid-wide reduction kernel."""
    rng = np.random.default_rng(len(einstr) + len(shapes))
    ops = [rng.uniform(0.5, 1.5, s).astype(dtype) for s in shapes]
    lst = planner.contract_path(einstr, *shapes, optimize=tuple([tuple(range(len(shapes)))]))[1]
    assert len(lst) == 1
    ref_T, ref_c = oracle.core_contract(ops, lst)
    T, c = gpu_engine.core_contract(ops, lst)
    assert_split_close(T, c, ref_T, ref_c, dtype == np.float32, what=einstr)
    # signed data: compare the dense value norm-wise
    ops = [rng.uniform(-1, 1, s).astype(dtype) for s in shapes]
    ref_T, ref_c = oracle.core_contract(ops, lst)
    T, c = gpu_engine.core_contract(ops, lst)
    ref = np.asarray(ref_T, dtype=np.float64) * np.exp(float(ref_c))
    got = np.asarray(T, dtype=np.float64) * np.exp(float(c))
    scale = np.sqrt(float(np.prod(shapes[0])))  # cancellation: error relative to the size of the terms
    tol = (1e-4 if dtype == np.float32 else 1e-11) * scale
    assert np.abs(got - ref).max() <= tol, (got, ref)
