"""The drop-in seams of SURVEY.md section 8b.

* Seam B (``install``): with the reference importable (this container: /root/reference through the
  ``opt_einsum`` stand-in of oracle/oe_shim.py) the UNMODIFIED ``contractn.contract`` / ``TN.contract``
  must reach ``contractn_b200.core_contract`` with the reference's own call convention
  (einsum.py:303) and hand its result on unchanged.  The arithmetic behind the seam needs a B200 and
  is covered by the ``-m gpu`` tests; here the device executor is replaced by the oracle so that the
  PLUMBING is what is under test (the reference's own ``tests/test_einsum.py:18-64`` logic runs
  through the patched seam).
* Seam A (``contractn_b200.backend``): per-op module, exercised on the GPU with the reference's
  executor loop restated over the module's functions.
"""
import os

import numpy as np
import pytest

from contractn_b200 import engine, planner
from oracle import contract_oracle as oracle

REF = "/root/reference"


@pytest.fixture()
def reference():
    if not os.path.isdir(os.path.join(REF, "contractn")):
        pytest.skip("the reference checkout is only present in the build container")
    from oracle import oe_shim

    return oe_shim.import_reference()


def test_install_patches_the_reference_executor(reference, monkeypatch):
    import contractn
    import contractn.einsum as ref_einsum

    calls = []

    def fake_run_numpy(arrays, contraction_list, out_mode=engine.CTN_OUT_SPLIT):
        # stands in for libctn_b200 (no GPU here): same contract as engine._run_numpy
        calls.append((len(arrays), len(contraction_list)))
        T, c = oracle.core_contract([np.asarray(a) for a in arrays], contraction_list)
        return np.asarray(T), np.array(c, dtype=np.float64)

    monkeypatch.setattr(engine, "_run_numpy", fake_run_numpy)
    original = engine.install(contractn)
    try:
        assert ref_einsum._core_contract is engine.core_contract
        # the reference's own hot-path test (tests/test_einsum.py:18-64): all-ones open-boundary MPS
        for n_cores, bdim, idim, split in [(2, 3, 2, True), (4, 5, 3, False), (6, 4, 2, True)]:
            tn = contractn.TN()
            cores = []
            for i in range(n_cores):
                shape = (bdim, idim) if i == 0 else (bdim, idim) if i == n_cores - 1 else (bdim, idim, bdim)
                cores.append(tn.add_dense_node(np.ones(shape)))
            for i in range(n_cores - 1):
                tn.connect_nodes(cores[i], cores[i + 1], 0 if i == 0 else 2, 0)
            out = tn.contract(split_format=split)
            log_expected = (n_cores - 1) * np.log(bdim)
            if split:
                T, c = out
                assert T.shape == (idim,) * n_cores
                assert np.allclose(np.log(T) + c, log_expected)
            else:
                assert out.shape == (idim,) * n_cores
                assert np.allclose(np.log(out), log_expected)
        # module-level entry with kwargs the reference forwards (einsum.py:271-305)
        a, b = np.random.default_rng(0).uniform(0.5, 1.5, (2, 4, 5))[0], np.random.default_rng(1).uniform(0.5, 1.5, (5, 6))
        T, c = contractn.contract("ab,bc->ac", a, b, split_format=True)
        np.testing.assert_allclose(T * np.exp(c), a @ b, rtol=1e-12)
        assert calls and all(n_ops >= 2 and n_steps == n_ops - 1 for n_ops, n_steps in calls)
    finally:
        ref_einsum._core_contract = original
    assert ref_einsum._core_contract is original


def _reference_loop(backend, operands, contraction_list):
    """einsum.py:341-391 restated over a backend MODULE (the functions the reference looks up by name)."""
    operands = list(operands)
    log_scale = backend.zeros(())
    for inds, idx_rm, einsum_str, _, blas_flag in contraction_list:
        tmp = [operands.pop(x) for x in inds]
        if blas_flag and "EINSUM" not in blas_flag:
            input_str, results_index = einsum_str.split("->")
            left, right = input_str.split(",")
            tensor_result = "".join(s for s in left + right if s not in idx_rm)
            lpos, rpos = zip(*sorted((left.find(s), right.find(s)) for s in idx_rm)) if idx_rm else ((), ())
            new = backend.tensordot(tmp[0], tmp[1], axes=(tuple(lpos), tuple(rpos)))
            if tensor_result != results_index:
                new = backend.transpose(new, tuple(tensor_result.index(s) for s in results_index))
        else:
            new = backend.einsum(einsum_str, *tmp)
        norm = backend.sum(backend.abs(new))  # stabilize, einsum.py:89-107
        rescale = norm / max(1, new.size)
        cond = norm > 1e-7
        new = backend.where(cond, new / rescale, new)
        log_scale = backend.where(cond, log_scale + backend.log(rescale), log_scale)
        operands.append(new)
    return operands[0], log_scale


@pytest.mark.gpu
def test_seam_a_backend_module_runs_the_reference_loop(gpu_engine):
    from contractn_b200 import backend

    rng = np.random.default_rng(3)
    einstr = "ab,acd,ce,bfg,dfhi,ehj,gk,ikl,jl->"
    shapes = [(4, 4), (4, 4, 4), (4, 4), (4, 4, 4), (4, 4, 4, 4), (4, 4, 4), (4, 4), (4, 4, 4), (4, 4)]
    ops = [rng.uniform(0, 1, s) for s in shapes]
    lst = planner.contract_path(einstr, *shapes)[1]
    T, c = _reference_loop(backend, ops, lst)
    want_T, want_c = oracle.core_contract(ops, lst)
    np.testing.assert_allclose(T, want_T, rtol=1e-11)
    assert abs(float(c) - float(want_c)) < 1e-9
    # the individual ops against numpy
    a, b = rng.uniform(0, 1, (6, 7, 8)), rng.uniform(0, 1, (8, 7, 5))
    np.testing.assert_allclose(backend.tensordot(a, b, axes=((1, 2), (1, 0))), np.tensordot(a, b, axes=((1, 2), (1, 0))), rtol=1e-12)
    np.testing.assert_allclose(backend.einsum("abc,cbd->ad", a, b), np.einsum("abc,cbd->ad", a, b), rtol=1e-12)
    x, y = rng.uniform(0, 1, (300, 2, 64)), rng.uniform(0, 1, (300, 2))
    np.testing.assert_allclose(backend.einsum("bps,bp->bs", x, y), np.einsum("bps,bp->bs", x, y), rtol=1e-12)
