"""Boundary-sharded sweep on the device plans, one GPU (world 1: the exchanges are identities, but the
segment plans, split-label views, output layouts and log-scale bookkeeping are the multi-GPU ones).
The N > 1 run is benchmarks/check_sharded_sweep.py under torchrun (bench.py's ``multi_gpu.sharded_sweep``)."""
import numpy as np
import pytest
import torch

from contractn_b200 import networks, planner, sharded
from oracle import contract_oracle as oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L,D,min_numel", [(4, 4, 64), (5, 4, 256), (6, 8, 1 << 12)])
def test_grid_matches_oracle(L, D, min_numel):
    tn, path = networks.peps_tn(L, D)
    einstr = tn.einsum_str
    ops = [torch.as_tensor(p).cuda() for p in tn.params]
    sw = sharded.ShardedSweep(einstr, ops, optimize=path, min_numel=min_numel)
    assert any(s.exchange == "a2a" for s in sw.segments)
    for _ in range(2):      # second run: bound plans and buffers are reused
        T, c = sw.run()
    lst = planner.cached_contract_list(einstr, tuple(p.shape for p in tn.params), optimize=path)
    To, co = oracle.contract(lst, *tn.params, split_format=True)
    np.testing.assert_allclose(T.cpu().numpy(), np.asarray(To), rtol=1e-12)
    assert abs(float(c[0]) - float(co)) < 1e-9


def test_open_output_and_batch_label():
    tn, path = networks.mps_tn(sites=8, bond=16, phys=2, batch=64, dtype=np.float64)
    einstr = tn.einsum_str
    ops = [torch.as_tensor(p).cuda() for p in tn.params]
    T, c = sharded.contract_sharded_sweep(einstr, *ops, optimize=path, min_numel=256)
    lst = planner.cached_contract_list(einstr, tuple(p.shape for p in tn.params), optimize=path)
    To, co = oracle.contract(lst, *tn.params, split_format=True)
    np.testing.assert_allclose(T.cpu().numpy(), np.asarray(To), rtol=1e-12)
    assert abs(float(c[0]) - float(co)) < 1e-9
