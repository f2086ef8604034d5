"""BASELINE configurations at their full / largest feasible tiers, against the ORACLE
on the same inputs where the CPU finishes in seconds, and through size-independent properties:

* cfg 3 literal (512 sites, bond 512, batch 4096, f32): the oracle in float64 (~20 s on the host): mantissas
  to rtol 1e-5 (north_star), log-scale to 5e-5 (510 tensor-core steps); plus batch independence and the
  ln 2 gauge shift (SURVEY finding 0.4).
* cfg 4 tier M (CP, I = R = 1024, f64, 8.6 GB output): the oracle needs minutes and 17 GB here, so sampled
  entries against the defining sum, and the canonical form mean|T-hat| = 1 (contractn/einsum.py:89-107).
* cfg 5 parity tier (8x8 grid, D = 8, f64): the oracle (~3 s): rtol 1e-12 / atol 1e-9 on the log-scale; the
  sum over the 8 values of one sliced bond equals the unsliced contraction; halving one site subtracts ln 2.
"""
import math

import numpy as np
import pytest

from contractn_b200 import planner
from oracle import contract_oracle as oracle

pytestmark = pytest.mark.gpu


def _device_ops(params):
    import torch

    return [torch.as_tensor(p).cuda() for p in params]


def test_cfg3_literal_batch_independence_and_gauge(gpu_engine):
    import torch
    from contractn_b200.networks import mps_tn

    sites, batch = 512, 4096
    tn, path = mps_tn(sites=sites, bond=512, batch=batch)
    ops = _device_ops(tn.params)
    einstr = tn.einsum_str
    shapes = tuple(tuple(o.shape) for o in ops)
    lst = planner.cached_contract_list(einstr, shapes, optimize=path)
    T, c = gpu_engine.core_contract(ops, lst)
    assert T.shape == (batch,) and T.dtype == torch.float32
    assert torch.isfinite(T).all() and (T > 0).all()
    assert abs(float(T.double().abs().mean()) - 1.0) < 1e-5
    # the float64 log-scale register of the device (the torch path hands back float32 like the reference's
    # torch backend; the plan keeps float64)
    plan = gpu_engine.get_plan(lst, shapes, np.float32, strides=tuple(tuple(o.stride()) for o in ops),
                               device=torch.cuda.current_device())
    _, tail = plan.execute_device(ops)
    c64 = float(tail[0].item())
    # ---- the oracle in float64 on the same inputs (einsum.py:326-393 restated; ~20 s of host time) ----
    want_T, want_c = oracle.core_contract([np.asarray(p, dtype=np.float64) for p in tn.params], lst)
    np.testing.assert_allclose(T.cpu().numpy().astype(np.float64), want_T, rtol=1e-5)
    assert abs(c64 - float(want_c)) < 5e-5, (c64, float(want_c))
    log_full = T.double().log() + c.double()

    half = batch // 2  # still a tcgen05-sized GEMM per site, so both runs share the kernel's rounding model
    sub = [o[:half] if i >= sites else o for i, o in enumerate(ops)]
    sub_shapes = tuple(tuple(o.shape) for o in sub)
    sub_lst = planner.cached_contract_list(einstr, sub_shapes, optimize=path)
    Ts, cs = gpu_engine.core_contract(sub, sub_lst)
    log_sub = Ts.double().log() + cs.double()
    # the torch float32 path returns the log-scale in float32 like the reference's torch backend does
    # (c ~ 3188: one float32 quantum is 2.4e-4), so the common offset is only good to ~1e-3 while the
    # per-sample spread of the difference is independent of it
    diff = log_full[:half] - log_sub
    assert float(diff.abs().max()) < 1e-3
    assert float(diff.max() - diff.min()) < 2e-5  # float32 path: 1022 steps of rounding

    ops2 = list(ops)
    ops2[100] = ops[100] * 2.0
    T2, c2 = gpu_engine.core_contract(ops2, lst)
    assert torch.equal(T2, T)
    assert abs(float(c2) - float(c) - math.log(2.0)) < 5e-4  # float32 quantum of the returned register at c ~ 3188
    _, tail2 = plan.execute_device(ops2)
    assert abs(float(tail2[0].item()) - c64 - math.log(2.0)) < 1e-6  # the device's float64 register


def test_cfg4_tier_m_sampled_entries(gpu_engine):
    import torch
    from contractn_b200.networks import cp_tn

    tn, optimize = cp_tn(I=1024, R=1024)
    params = tn.params
    ops = _device_ops(params)
    einstr = tn.einsum_str
    shapes = tuple(tuple(o.shape) for o in ops)
    lst = planner.cached_contract_list(einstr, shapes, optimize=optimize)
    T, c = gpu_engine.core_contract(ops, lst)
    assert T.shape == (1024, 1024, 1024) and T.dtype == torch.float64
    assert abs(float(T.abs().mean()) - 1.0) < 1e-12
    rng = np.random.default_rng(4)
    idx = rng.integers(0, 1024, size=(2000, 3))
    got = T[idx[:, 0], idx[:, 1], idx[:, 2]].cpu().numpy() * math.exp(float(c))
    f1, f2, f3 = params
    want = np.einsum("na,na,na->n", f1[idx[:, 0]], f2[idx[:, 1]], f3[idx[:, 2]])
    np.testing.assert_allclose(got, want, rtol=1e-12)
    del T
    torch.cuda.empty_cache()


def test_cfg5_parity_tier_slicing_and_gauge(gpu_engine):
    from contractn_b200.networks import peps_tn
    from contractn_b200 import parallel

    tn, path = peps_tn(8, 8)
    ops = _device_ops(tn.params)
    einstr = tn.einsum_str
    shapes = tuple(tuple(o.shape) for o in ops)
    lst = planner.cached_contract_list(einstr, shapes, optimize=path)
    T, c = gpu_engine.core_contract(ops, lst)
    assert T.numel() == 1 and abs(float(T) - 1.0) < 1e-12   # scalar: the mantissa is exactly the sign
    # ---- the oracle on the same inputs (6.5e11 FLOP, ~3 s of host time) ----
    want_T, want_c = oracle.core_contract(tn.params, lst)
    np.testing.assert_allclose(float(T), float(want_T), rtol=1e-12)
    assert abs(float(c) - float(want_c)) < 1e-9, (float(c), float(want_c))
    labels = parallel.choose_slice_labels(einstr, shapes, lst, 8)
    Ts, cs = parallel.contract_sliced(einstr, *ops, slice_labels=labels, optimize=path, split_format=True)
    assert abs(float(Ts) - 1.0) < 1e-12
    assert abs(float(cs) - float(c)) < 1e-9 * max(1.0, abs(float(c)))
    ops2 = list(ops)
    ops2[27] = ops[27] * 0.5
    T2, c2 = gpu_engine.core_contract(ops2, lst)
    assert abs(float(c2) - float(c) + math.log(2.0)) < 1e-9
