"""Complex operands through the device engine (real kernels, complex_lowering.py) against the oracle, which takes
complex128 through numpy exactly as the reference does (abs -> real scale, contractn/einsum.py:95-106)."""
import numpy as np
import pytest

from oracle import contract_oracle as oracle
from contractn_b200 import planner

pytestmark = pytest.mark.gpu

CASES = [
    ("ab,bc->ac", [(30, 40), (40, 50)]),
    ("ab,bc,cd->ad", [(130, 140), (140, 150), (150, 160)]),   # GEMM-class real steps with a doubled K
    ("ab,ab,ab->ab", [(64, 48)] * 3),
    ("bi,bj,ij->b", [(600, 30), (600, 40), (30, 40)]),
    ("ab,bc->", [(30, 40), (40, 50)]),
    ("abc->ca", [(20, 30, 40)]),
]


@pytest.mark.parametrize("einstr,shapes", CASES)
@pytest.mark.parametrize("cdt", [np.complex128, np.complex64])
def test_complex_numpy_and_torch(gpu_engine, einstr, shapes, cdt):
    import torch

    rng = np.random.default_rng(6)
    ops = [(rng.normal(size=s) + 1j * rng.normal(size=s)).astype(cdt) for s in shapes]
    lst = planner.contract_path(einstr, *shapes)[1]
    want_T, want_c = oracle.core_contract([o.astype(np.complex128) for o in ops], lst)
    rtol = 1e-10 if cdt == np.complex128 else 2e-4
    scale = np.abs(want_T).max()

    T, c = gpu_engine.core_contract(ops, lst)
    assert T.dtype == cdt and T.shape == np.asarray(want_T).shape
    assert np.abs(T - want_T).max() <= rtol * scale and abs(float(c) - float(want_c)) <= rtol * 10

    dense = gpu_engine.contract(einstr, *ops)
    want_dense = want_T * np.exp(want_c)
    assert np.abs(dense - want_dense).max() <= rtol * np.abs(want_dense).max()

    tT, tc = gpu_engine.core_contract([torch.as_tensor(o).cuda() for o in ops], lst)
    assert tT.is_cuda and tT.is_complex()
    assert np.abs(tT.cpu().numpy() - want_T).max() <= rtol * scale and abs(float(tc) - float(want_c)) <= rtol * 10


def test_real_and_complex_operands_mix(gpu_engine):
    rng = np.random.default_rng(8)
    a = rng.normal(size=(20, 30))
    b = rng.normal(size=(30, 10)) + 1j * rng.normal(size=(30, 10))
    lst = planner.contract_path("ab,bc->ac", a.shape, b.shape)[1]
    want_T, want_c = oracle.core_contract([a.astype(np.complex128), b], lst)
    T, c = gpu_engine.core_contract([a, b], lst)
    assert np.abs(T - want_T).max() <= 1e-10 * np.abs(want_T).max() and abs(float(c) - float(want_c)) < 1e-9
