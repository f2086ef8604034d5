"""Randomised parity: random einsum networks (2-5 operands, random extents, hyperedges, batch and
summed indices, strided / transposed / broadcast views) through every optimizer, float64 and
float32, against the CPU oracle executing the SAME contraction list."""
import numpy as np
import pytest

from oracle import contract_oracle as oracle
from contractn_b200 import planner

pytestmark = pytest.mark.gpu

SYMS = "abcdefgh"


def random_network(rng, big=False):
    n_ops = int(rng.integers(2, 6))
    n_sym = int(rng.integers(2, 7))
    syms = list(SYMS[:n_sym])
    hi = 40 if big else 7
    dims = {s: int(rng.integers(1, hi)) for s in syms}
    terms = []
    for _ in range(n_ops):
        k = int(rng.integers(1, min(4, n_sym) + 1))
        terms.append("".join(rng.choice(syms, size=k, replace=False)))
    flat = "".join(terms)
    used = sorted(set(flat))
    # output: a random subset of the used symbols, in random order (hyperedges included)
    n_out = int(rng.integers(0, len(used) + 1))
    out = "".join(rng.permutation(used)[:n_out])
    return ",".join(terms) + "->" + out, [tuple(dims[c] for c in t) for t in terms]


def random_view(rng, shape, dtype):
    """An array of the given shape that is a non-trivial view: transposed storage, strided
    slices or an offset base pointer."""
    kind = int(rng.integers(0, 4))
    if kind == 0 or len(shape) == 0:
        return rng.uniform(0.2, 1.0, shape).astype(dtype)
    if kind == 1:  # Fortran / permuted storage
        perm = rng.permutation(len(shape))
        base = rng.uniform(0.2, 1.0, tuple(shape[p] for p in perm)).astype(dtype)
        return base.transpose(np.argsort(perm))
    if kind == 2:  # every other element along each axis
        base = rng.uniform(0.2, 1.0, tuple(2 * d + 1 for d in shape)).astype(dtype)
        return base[tuple(slice(1, 2 * d + 1, 2) for d in shape)]
    flat = rng.uniform(0.2, 1.0, int(np.prod(shape)) + 3).astype(dtype)
    return flat[3:].reshape(shape)  # odd offset: not 16-byte aligned


@pytest.mark.parametrize("seed", range(40))
def test_random_networks_f64(gpu_engine, seed):
    rng = np.random.default_rng(1000 + seed)
    einstr, shapes = random_network(rng, big=seed % 4 == 0)
    ops = [random_view(rng, s, np.float64) for s in shapes]
    for opt in ("auto", "greedy"):
        lst = planner.contract_path(einstr, *shapes, optimize=opt)[1]
        want_T, want_c = oracle.core_contract(ops, lst)
        T, c = gpu_engine.core_contract(ops, lst)
        assert T.shape == want_T.shape, einstr
        np.testing.assert_allclose(T, want_T, rtol=1e-11, atol=0, err_msg=f"{einstr} {shapes} {opt}")
        assert abs(float(c) - float(want_c)) < 1e-9, f"{einstr} {shapes} {opt}"


@pytest.mark.parametrize("seed", range(20))
def test_random_networks_f32(gpu_engine, seed):
    rng = np.random.default_rng(5000 + seed)
    einstr, shapes = random_network(rng, big=seed % 3 == 0)
    ops = [random_view(rng, s, np.float32) for s in shapes]
    lst = planner.contract_path(einstr, *shapes)[1]
    want_T, want_c = oracle.core_contract([o.astype(np.float64) for o in ops], lst)
    T, c = gpu_engine.core_contract(ops, lst)
    assert T.dtype == np.float32 and T.shape == want_T.shape
    np.testing.assert_allclose(T.astype(np.float64), want_T, rtol=2e-5, atol=0, err_msg=f"{einstr} {shapes}")
    assert abs(float(c) - float(want_c)) < 1e-4


def test_weight_sharing_and_input_nodes(gpu_engine):
    """Clone nodes share one parameter (the reference asserts here);
    input nodes are filled at call time (ctn.py:167)."""
    from contractn_b200 import TN

    rng = np.random.default_rng(9)
    A = rng.uniform(0.1, 1, (3, 5, 5))
    tn = TN()
    first = tn.add_dense_node(A)
    nodes = [first] + [tn.add_duplicate_node(first) for _ in range(3)]
    x_nodes = [tn.add_input_node((3,)) for _ in range(4)]
    for i in range(3):
        tn.connect_nodes(nodes[i], nodes[i + 1], 2, 1)
    for n, x in zip(nodes, x_nodes):
        tn.connect_nodes(n, x, 0, 0)
    xs = [rng.uniform(0.1, 1, (3,)) for _ in range(4)]
    fun = tn.make_contract_fun(split_format=False)
    got = fun(tn.params, xs)
    mats = [np.einsum("pab,p->ab", A, x) for x in xs]
    want = mats[0] @ mats[1] @ mats[2] @ mats[3]
    np.testing.assert_allclose(got, want, rtol=1e-12)
    assert len(tn.params) == 1
