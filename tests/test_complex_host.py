"""Complex operands lowered to real contractions (contractn_b200/complex_lowering.py), checked on the CPU: the engine
call is replaced by a numpy stand-in that returns every step in an arbitrary gauge, so what is tested is the lowering
-- the structure tensor, the per-step real einsum strings, the final re-normalisation to the reference's canonical
form (complex modulus, contractn/einsum.py:95-106) -- against the oracle on complex128 inputs."""
import math

import numpy as np
import pytest

from contractn_b200 import complex_lowering as cl, planner
from oracle import contract_oracle as oracle

CASES = [
    ("ab,bc->ac", [(3, 4), (4, 5)]),
    ("ab,bc,cd->ad", [(3, 4), (4, 5), (5, 6)]),
    ("ab,ab,ab->ab", [(3, 4)] * 3),
    ("bi,bj,ij->b", [(6, 3), (6, 4), (3, 4)]),
    ("ab,bc->", [(3, 4), (4, 5)]),
    ("abc->ca", [(2, 3, 4)]),
]


def _stand_in(einstr, *ops, optimize=None):
    import string

    table = {}  # (np.einsum takes ASCII letters only; the engine takes any symbol)
    for ch in einstr:
        if ch not in ",->" and ch not in table:
            table[ch] = string.ascii_letters[len(table)]
    t = np.einsum("".join(table.get(ch, ch) for ch in einstr), *ops)
    kappa = np.abs(t).mean() * 3.7  # (off the engine's gauge on purpose)
    return t / kappa, math.log(kappa)


def _to_real(x):
    x = np.ascontiguousarray(x, dtype=np.complex128)
    return x.view(np.float64).reshape(x.shape + (2,))


def _to_complex(x):
    x = np.ascontiguousarray(x)
    return x.view(np.complex128).reshape(x.shape[:-1])


@pytest.mark.parametrize("einstr,shapes", CASES)
def test_lowered_contraction_matches_the_oracle(einstr, shapes):
    rng = np.random.default_rng(2)
    ops = [rng.normal(size=s) + 1j * rng.normal(size=s) for s in shapes]
    lst = planner.contract_path(einstr, *shapes)[1]
    want_T, want_c = oracle.core_contract(ops, lst)
    got_T, got_c = cl.run(ops, lst, _stand_in, _to_real, cl.structure_tensor(np.float64), _to_complex,
                          lambda x: (float(np.abs(x).sum()), float(np.abs(x).mean())), math.log)
    assert got_T.shape == np.asarray(want_T).shape
    np.testing.assert_allclose(got_T, want_T, rtol=1e-11, atol=1e-13)
    assert abs(got_c - float(want_c)) < 1e-10


def test_structure_tensor_is_complex_multiplication():
    s = cl.structure_tensor(np.float64)
    a, b = 1.5 - 2.0j, -0.25 + 3.0j
    r = np.einsum("p,pqr,q->r", [a.real, a.imag], s, [b.real, b.imag])
    assert np.allclose(r, [(a * b).real, (a * b).imag])


def test_lowered_step_strings():
    real_str, path, with_s = cl.lower_step("ab,bc->ac")
    ins, out = real_str.split("->")
    assert with_s and path == ((0, 1), (0, 1)) and len(ins.split(",")) == 3
    assert ins.split(",")[0][:2] == "ab" and out[:2] == "ac" and len(out) == 3
    with pytest.raises(NotImplementedError):
        cl.lower_step("ab,bc,cd->ad")
