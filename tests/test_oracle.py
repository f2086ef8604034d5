"""The CPU oracle against the reference's golden vectors (no GPU needed).

The fixtures were produced by the UNMODIFIED reference (oracle/make_golden.py);
the oracle is the numpy restatement the GPU parity tests compare against.
"""
import numpy as np
import pytest

from oracle import contract_oracle as oracle
from contractn_b200 import planner

from _cases import Golden, golden_names


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference(name):
    g = Golden(name)
    T, c = oracle.core_contract(g.operands, g.contraction_list)
    # same numpy calls in the same order as the reference: agreement to rounding
    # (bit-exact on the generating machine; BLAS threading may differ elsewhere)
    np.testing.assert_allclose(T, g.ref_T, rtol=1e-13, atol=0)
    np.testing.assert_allclose(c, g.ref_c, rtol=0, atol=1e-12)
    assert T.dtype == g.ref_T.dtype
    assert np.asarray(c).dtype == np.float64  # numpy keeps a float64 register even for f32 data
    with np.errstate(over="ignore"):
        dense = oracle.destabilize(T, c)
    np.testing.assert_allclose(dense, g.ref_dense, rtol=1e-13, atol=0)


def test_readme_known_answers():
    """README.md:33 and README.md:74-76."""
    g = Golden("readme_hub")
    np.testing.assert_allclose(g.ref_dense, [1.0, 0.36603234], rtol=1e-8)
    T, c = oracle.core_contract(g.operands, g.contraction_list)
    np.testing.assert_allclose(oracle.destabilize(T, c), [1.0, 0.99 ** 100], rtol=1e-12)
    g = Golden("readme_chain")
    T, c = oracle.core_contract(g.operands, g.contraction_list)
    np.testing.assert_allclose(T, [1.0, 1.0, 1.0], rtol=1e-12)
    np.testing.assert_allclose(c, 1000 * np.log(3.0), rtol=1e-13)
    assert abs(float(c) - 1098.61228867) < 1e-8
    with np.errstate(over="ignore"):
        assert np.all(np.isinf(oracle.destabilize(T, c)))  # "[inf inf inf]"


@pytest.mark.parametrize("name", [n for n in golden_names() if n not in ("zero_operand", "tiny_norm_skip")])
def test_split_output_is_canonical(name):
    """SURVEY.md finding 0.4: (T, c) == (X/mean|X|, ln mean|X|) for the plain
    einsum X, whenever the last step is rescaled -- path independent."""
    g = Golden(name)
    if len(g.operands) > 40:
        pytest.skip("dense einsum of this many operands is not practical")
    ops = [np.asarray(o, dtype=np.float64) for o in g.operands]
    X = oracle._einsum_az(g.einstr, *ops, optimize="greedy")
    T, c = oracle.canonical_split(X)
    rtol = 1e-5 if g.is_f32 else 1e-10
    if g.signed:
        assert np.linalg.norm((T - g.ref_T).ravel()) <= rtol * 100 * np.linalg.norm(T.ravel())
    else:
        np.testing.assert_allclose(g.ref_T, T, rtol=rtol)
    np.testing.assert_allclose(g.ref_c, c, atol=1e-4 if g.is_f32 else 1e-9)


def test_stabilize_rules():
    """einsum.py:89-107: mean-abs, threshold on the SUM, sign kept."""
    t = np.array([[1.0, -3.0], [0.0, 4.0]])
    out, c = oracle.stabilize(t, np.zeros(()))
    np.testing.assert_allclose(out, t / 2.0)
    np.testing.assert_allclose(c, np.log(2.0))
    tiny = np.full((10,), 1e-9)  # sum = 1e-8 <= 1e-7 -> untouched
    out, c = oracle.stabilize(tiny, np.zeros(()))
    assert np.array_equal(out, tiny) and float(c) == 0.0
    edge = np.full((10,), 2e-8)  # sum = 2e-7 > 1e-7 although the mean is 2e-8
    out, c = oracle.stabilize(edge, np.zeros(()))
    np.testing.assert_allclose(out, np.ones(10))
    zero = np.zeros((3, 3))
    out, c = oracle.stabilize(zero, np.array(1.5))
    assert np.array_equal(out, zero) and float(c) == 1.5


def test_oracle_executes_planner_lists():
    rng = np.random.default_rng(0)
    ops = [rng.uniform(0, 1, s) for s in [(4, 5), (5, 6, 3), (3, 6, 2), (2, 4)]]
    for opt in ("greedy", "optimal", "branch-all", "branch-2", "branch-1", "auto"):
        lst = planner.contract_path("ab,bcd,dce,ea->", *[o.shape for o in ops], optimize=opt)[1]
        T, c = oracle.core_contract(ops, lst)
        np.testing.assert_allclose(T * np.exp(c), np.einsum("ab,bcd,dce,ea->", *ops), rtol=1e-12)
