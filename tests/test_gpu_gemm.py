"""The strided DMMA GEMM kernel against numpy on shapes that exercise every
staging layout (K-fast / MN-fast / element-wise), tile configuration, batch,
multi-mode groups, ragged edges and split-K.  float64, rtol 1e-12."""
import numpy as np
import pytest

from oracle import contract_oracle as oracle
from contractn_b200 import planner

pytestmark = pytest.mark.gpu

CASES = [
    # (einsum, shapes, what)
    ("ab,bc->ac", [(256, 128), (128, 192)], "plain NN: A K-fast, B N-fast"),
    ("ab,cb->ac", [(200, 64), (150, 64)], "NT: both K-fast"),
    ("ba,bc->ac", [(64, 200), (64, 150)], "TN: both MN-fast"),
    ("ba,cb->ac", [(96, 130), (70, 96)], "TT"),
    ("ab,bc->ca", [(130, 40), (40, 70)], "transposed output"),
    ("ab,bc->ac", [(129, 67), (67, 131)], "odd sizes: element-wise staging, ragged tiles"),
    ("ab,bc->ac", [(1000, 24), (24, 16)], "narrow N: 128x32 tiles"),
    ("ab,bc->ac", [(16, 24), (24, 1000)], "narrow M: 32x128 tiles"),
    ("ab,b->a", [(512, 300), (300,)], "GEMV"),
    ("a,ab->b", [(300,), (300, 512)], "vector-matrix"),
    ("ab,ab->", [(300, 400), (300, 400)], "DOT: reduction kernel"),
    ("ak,bk->ab", [(8, 120000), (8, 120000)], "8x8 output, K = 120000: split-K"),
    ("ab,ba->", [(300, 400), (400, 300)], "transposed DOT"),
    ("abc,bcd->ad", [(40, 30, 20), (30, 20, 50)], "two contiguous summed modes fuse"),
    ("abc,cbd->ad", [(40, 30, 20), (20, 30, 50)], "two summed modes, not fusable in B"),
    ("bij,bjk->bik", [(6, 50, 70), (6, 70, 40)], "batched"),
    ("ibj,jbk->bik", [(50, 6, 70), (70, 6, 40)], "batch index in the middle"),
    ("br,prs->bps", [(300, 64), (2, 64, 64)], "cfg 3 step (TDOT)"),
    ("dbc,eb->dce", [(20, 64, 30), (50, 64)], "cfg 4 Tucker step (TDOT, summed mode mid-tensor)"),
    ("cab,da->bcd", [(30, 64, 40), (50, 64)], "cfg 4 CP step"),
    ("abcdefgh,efxy->abcdghxy", [(4,) * 8, (4, 4, 8, 8)], "cfg 5 boundary step: summed legs mid-tensor"),
    ("abcd,cdef->abef", [(8, 8, 8, 8), (8, 8, 8, 8)], "PEPS block"),
    ("ab,bc->ac", [(40, 5000), (5000, 30)], "small output, long K: split-K"),
    ("aab,bc->ac", [(20, 20, 64), (64, 30)], "diagonal of an operand"),
    ("ab,bc->ac", [(64, 1), (300, 50)], "broadcast summed index -> not a GEMM, generic"),
    ("ab,bc->ac", [(4100, 40), (40, 24)], "skinny: rows from the left operand"),
    ("ab,bc->ac", [(24, 40), (40, 4100)], "skinny: rows from the right operand"),
    ("abcdef,efxy->abcdxy", [(6, 7, 8, 9, 4, 5), (4, 5, 3, 2)], "skinny: PEPS absorb, summed legs last"),
    ("abefcd,efxy->abxycd", [(9, 8, 4, 5, 7, 6), (4, 5, 3, 2)], "skinny: summed legs mid-tensor, output interleaved"),
    ("xe,abcde->abcdx", [(5, 6), (8, 8, 8, 8, 6)], "skinny swapped, single summed leg"),
    ("ab,bc->ac", [(3000, 7), (7, 64)], "skinny: N = 64 accumulators"),
]


@pytest.mark.parametrize("einstr,shapes,what", CASES, ids=[c[2] for c in CASES])
def test_gemm_step(gpu_engine, einstr, shapes, what):
    rng = np.random.default_rng(hash(einstr) % (2 ** 32))
    ops = [rng.uniform(0.1, 1.0, s) for s in shapes]
    lst = planner.contract_path(einstr, *shapes)[1]
    T, c = gpu_engine.core_contract(ops, lst)
    want_T, want_c = oracle.core_contract(ops, lst)
    np.testing.assert_allclose(T, want_T, rtol=1e-12, err_msg=what)
    assert abs(float(c) - float(want_c)) < 1e-9


def test_gemm_kinds_are_selected(gpu_engine):
    def kinds(einstr, shapes):
        lst = planner.contract_path(einstr, *shapes)[1]
        return gpu_engine.Plan(lst, shapes, np.float64).launch_kinds

    assert kinds("ab,bc->ac", [(256, 128), (128, 192)]) == ("gemm", "finalize")
    assert kinds("ab,ab->", [(300, 400)] * 2) == ("generic", "finalize")  # DOT class: grid-wide reduction kernel
    assert kinds("ak,bk->ab", [(8, 120000)] * 2) == ("gemm", "gemm_reduce", "finalize")  # small output, long K
    assert kinds("bps,bp->bs", [(300, 2, 64), (300, 2)]) == ("generic", "finalize")
    assert kinds("ab,bc->ac", [(4100, 40), (40, 24)]) == ("gemm", "finalize")  # skinny variant


def test_gemm_on_strided_views_and_misaligned_pointers(gpu_engine):
    rng = np.random.default_rng(4)
    big_a = rng.uniform(0.1, 1, (300, 260))
    big_b = rng.uniform(0.1, 1, (270, 310))
    a = big_a[3:259:2, 1:201]       # row stride 520, odd element offset -> 8-byte staging
    b = big_b[5:205, 7:247:3]       # column stride 3
    lst = planner.contract_path("ab,bc->ac", a.shape, b.shape)[1]
    T, c = gpu_engine.core_contract([a, b], lst)
    want_T, want_c = oracle.core_contract([a, b], lst)
    np.testing.assert_allclose(T, want_T, rtol=1e-12)
    assert abs(float(c) - float(want_c)) < 1e-9
    import torch

    ta = torch.as_tensor(np.ascontiguousarray(a)).cuda()
    flat = torch.empty(ta.numel() + 1, dtype=torch.float64, device="cuda")
    mis = flat[1:].view(ta.shape)  # 8-byte but not 16-byte aligned base
    mis.copy_(ta)
    tb = torch.as_tensor(np.ascontiguousarray(b)).cuda()
    T2, c2 = gpu_engine.core_contract([mis, tb], lst)
    np.testing.assert_allclose(T2.cpu().numpy(), want_T, rtol=1e-12)


def test_gemm_signed_data_normwise(gpu_engine):
    rng = np.random.default_rng(8)
    a, b = rng.normal(size=(300, 700)), rng.normal(size=(700, 200))
    lst = planner.contract_path("ab,bc->ac", a.shape, b.shape)[1]
    T, c = gpu_engine.core_contract([a, b], lst)
    want_T, want_c = oracle.core_contract([a, b], lst)
    assert np.linalg.norm(T - want_T) <= 1e-12 * np.linalg.norm(want_T)
    assert abs(float(c) - float(want_c)) < 1e-9


def test_peps_6x6_d4_against_oracle(gpu_engine):
    """cfg 5 in small: closed 6x6 grid, D = 4, greedy path (GEMM + TDOT steps)."""
    rng = np.random.default_rng(1005)
    L, D = 6, 4
    sym, terms, shapes = {}, [], []

    def s(key):
        if key not in sym:
            sym[key] = planner.get_symbol(len(sym))
        return sym[key]

    for r in range(L):
        for c in range(L):
            t = ""
            if r > 0: t += s(("v", r - 1, c))
            if c > 0: t += s(("h", r, c - 1))
            if r < L - 1: t += s(("v", r, c))
            if c < L - 1: t += s(("h", r, c))
            terms.append(t)
            shapes.append((D,) * len(t))
    einstr = ",".join(terms) + "->"
    ops = [rng.uniform(0, 1, sh) for sh in shapes]
    lst = planner.contract_path(einstr, *shapes, optimize="greedy")[1]
    T, c = gpu_engine.core_contract(ops, lst)
    want_T, want_c = oracle.core_contract(ops, lst)
    np.testing.assert_allclose(T, want_T, rtol=1e-12)
    assert abs(float(c) - float(want_c)) < 1e-9


# ---------------------------------------------------------------------------
# float32: tcgen05 3xTF32 kernel (M, N >= 16) and generic kernel otherwise
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("einstr,shapes,what", CASES, ids=[c[2] for c in CASES])
def test_gemm_step_f32(gpu_engine, einstr, shapes, what):
    """north_star tolerance for float32 values: rtol 1e-5 (against float64 truth)."""
    rng = np.random.default_rng(hash(einstr) % (2 ** 32))
    ops = [rng.uniform(0.1, 1.0, s).astype(np.float32) for s in shapes]
    lst = planner.contract_path(einstr, *shapes)[1]
    T, c = gpu_engine.core_contract(ops, lst)
    assert T.dtype == np.float32
    ops64 = [o.astype(np.float64) for o in ops]
    want_T, want_c = oracle.core_contract(ops64, lst)
    np.testing.assert_allclose(T.astype(np.float64), want_T, rtol=1e-5, err_msg=what)
    # log-scale: the tensor core truncates when adding into its FP32 accumulator (~6e-8 per chained
    # add, common to all entries); the kernel measures that bias on sampled entries and removes it
    # from the log-scale, so one step is good to float32 rounding
    assert abs(float(c) - float(want_c)) < 2e-6


def test_tf32_kernel_is_selected_and_accurate_for_long_k(gpu_engine):
    """Accumulation-chain control: K = 8192 still meets rtol 1e-5 on positive data."""
    rng = np.random.default_rng(12)
    a = rng.uniform(0.0, 1.0, (256, 8192)).astype(np.float32)
    b = rng.uniform(0.0, 1.0, (8192, 192)).astype(np.float32)
    lst = planner.contract_path("ab,bc->ac", a.shape, b.shape)[1]
    plan = gpu_engine.Plan(lst, [a.shape, b.shape], np.float32)
    assert plan.launch_kinds == ("gemm", "finalize")
    T, c, _ = plan.execute_host([a, b])
    want = a.astype(np.float64) @ b.astype(np.float64)
    got = T.astype(np.float64) * np.exp(c)
    assert np.abs(got / want - 1).max() < 3e-5
    # the normalised mantissa (what split format returns) is tighter: the
    # accumulator's truncation bias is common to all entries and cancels
    wT, wc = oracle.canonical_split(want)
    np.testing.assert_allclose(T.astype(np.float64), wT, rtol=1e-5)


def test_tf32_signed_data_normwise(gpu_engine):
    rng = np.random.default_rng(13)
    a = rng.normal(size=(300, 700)).astype(np.float32)
    b = rng.normal(size=(700, 200)).astype(np.float32)
    lst = planner.contract_path("ab,bc->ac", a.shape, b.shape)[1]
    T, c = gpu_engine.core_contract([a, b], lst)
    want = a.astype(np.float64) @ b.astype(np.float64)
    got = T.astype(np.float64) * np.exp(float(c))
    assert np.linalg.norm(got - want) <= 1e-5 * np.linalg.norm(want)


# ---------------------------------------------------------------------------
# consumer fusion (float32): GEMM + hyperedge reduce in one tcgen05 kernel
# ---------------------------------------------------------------------------
FUSED_CASES = [
    ("br,prs,bp->bs", [(300, 64), (2, 64, 64), (300, 2)], ((0, 1), (0, 1)), "cfg 3 site, pr = 2"),
    ("br,prs,bp->bs", [(200, 96), (4, 96, 64), (200, 4)], ((0, 1), (0, 1)), "pr = 4"),
    ("prs,br,bp->bs", [(2, 40, 128), (260, 40), (260, 2)], ((0, 1), (0, 1)), "operands swapped"),
    ("br,prs,bp->bs", [(129, 72), (2, 72, 70), (129, 2)], ((0, 1), (0, 1)), "ragged rows / columns, element-wise B staging"),
    ("br,rps,pb->sb", [(256, 64), (64, 2, 128), (2, 256)], ((0, 1), (0, 1)), "r-major core, transposed weights and output"),
]


@pytest.mark.parametrize("einstr,shapes,path,what", FUSED_CASES, ids=[c[3] for c in FUSED_CASES])
def test_fused_consumer_f32(gpu_engine, einstr, shapes, path, what):
    rng = np.random.default_rng(hash(what) % (2 ** 32))
    ops = [rng.uniform(0.1, 1.0, s).astype(np.float32) for s in shapes]
    lst = planner.contract_path(einstr, *shapes, optimize=path)[1]
    plan = gpu_engine.Plan(lst, shapes, np.float32)
    assert plan.launch_kinds == ("gemm", "finalize"), plan.launch_kinds  # two steps, one kernel
    T, c, fl = plan.execute_host(ops)
    assert fl == 0
    want_T, want_c = oracle.core_contract([o.astype(np.float64) for o in ops], lst)
    np.testing.assert_allclose(T.astype(np.float64), want_T, rtol=1e-5, err_msg=what)
    assert abs(float(c) - float(want_c)) < 2e-6
    # and the unfused plan agrees
    exact = gpu_engine.Plan(lst, shapes, np.float32, flags=gpu_engine.CTN_PLAN_EXACT_STEPS)
    assert exact.launch_kinds == ("gemm", "generic", "finalize")
    T2, c2, _ = exact.execute_host(ops)
    np.testing.assert_allclose(T, T2, rtol=1e-5)


def test_fused_consumer_degenerate_falls_back(gpu_engine):
    """Zero weights: the reference keeps the log-scale of the GEMM step and leaves the zero
    result unscaled; the fused kernel flags it and the engine re-runs step by step."""
    rng = np.random.default_rng(3)
    shapes = [(300, 64), (2, 64, 64), (300, 2)]
    ops = [rng.uniform(0.1, 1.0, s).astype(np.float32) for s in shapes]
    ops[2][...] = 0.0
    lst = planner.contract_path("br,prs,bp->bs", *shapes, optimize=((0, 1), (0, 1)))[1]
    plan = gpu_engine.Plan(lst, shapes, np.float32)
    _, _, fl = plan.execute_host(ops)
    assert fl & gpu_engine.CTN_RESULT_DEGENERATE
    T, c = gpu_engine.core_contract(ops, lst)
    want_T, want_c = oracle.core_contract(ops, lst)
    assert np.array_equal(T, want_T)
    assert abs(float(c) - float(want_c)) < 1e-4
