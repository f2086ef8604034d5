"""Multi-rank host logic on CPU: world_size-2 gloo process groups.

The per-shard / per-slice contraction is done by the oracle here (the product
path needs a GPU); what is under test is the partitioning and the log-sum-exp
merges of (mantissa, log-scale) pairs in contractn_b200.parallel.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from contractn_b200 import parallel, planner
from oracle import contract_oracle as oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_contract(einstr, *ops, optimize="auto", split_format=True):
    ops = [np.asarray(o) for o in ops]
    lst = planner.contract_path(einstr, *[o.shape for o in ops], optimize=optimize)[1]
    return oracle.contract(lst, *ops, split_format=split_format)


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _batch_case(rank, world):
    rng = np.random.default_rng(11)
    n, B = 12, 64
    ops = [rng.uniform(0.2, 1.8, (3, B)) * (10.0 ** rng.integers(-3, 4)) for _ in range(n)]
    # make the shards very different in magnitude: per-rank scales differ a lot
    for o in ops[:4]:
        o[:, B // 2:] *= 1e-6
    einstr = ",".join(["ab"] * n) + "->ab"
    lo, hi = rank * B // world, (rank + 1) * B // world
    T, c = _oracle_contract(einstr, *[o[:, lo:hi] for o in ops])
    T, c = parallel.merge_batch_shards(torch.as_tensor(T), torch.as_tensor(c).reshape(1))
    full_T, full_c = _oracle_contract(einstr, *ops)
    return T.numpy(), float(c[0]), full_T[:, lo:hi], float(full_c)


def test_batch_shard_merge_restores_global_scale():
    for T, c, want_T, want_c in _run(_batch_case):
        np.testing.assert_allclose(T, want_T, rtol=1e-12)
        assert abs(c - want_c) < 1e-9


def _slice_case(rank, world):
    rng = np.random.default_rng(5)
    # 3x3 closed grid, D = 4
    einstr = "ab,acd,ce,bfg,dfhi,ehj,gk,ikl,jl->"
    shapes = [(4, 4), (4, 4, 4), (4, 4), (4, 4, 4), (4, 4, 4, 4), (4, 4, 4), (4, 4), (4, 4, 4), (4, 4)]
    ops = [torch.as_tensor(rng.uniform(-0.5, 1.0, s)) for s in shapes]
    T, c = parallel.contract_sliced(einstr, *ops, min_slices=8, contract_fn=_oracle_contract)
    want_T, want_c = _oracle_contract(einstr, *[o.numpy() for o in ops])
    T2, c2 = parallel.contract_sliced(einstr, *ops, slice_labels=("d",), contract_fn=_oracle_contract)
    return T.numpy(), float(c[0]), T2.numpy(), float(c2[0]), want_T, float(want_c)


def test_bond_slicing_sums_to_the_unsliced_result():
    for T, c, T2, c2, want_T, want_c in _run(_slice_case):
        np.testing.assert_allclose(T, want_T, rtol=1e-10)
        assert abs(c - want_c) < 1e-9
        np.testing.assert_allclose(T2, want_T, rtol=1e-10)
        assert abs(c2 - want_c) < 1e-9


def _open_slice_case(rank, world):
    rng = np.random.default_rng(6)
    einstr = "abc,bd,cde->ae"
    ops = [torch.as_tensor(rng.uniform(0, 1, s)) for s in [(5, 6, 7), (6, 8), (7, 8, 3)]]
    T, c = parallel.contract_sliced(einstr, *ops, slice_labels=("d", "b"), contract_fn=_oracle_contract)
    want_T, want_c = _oracle_contract(einstr, *[o.numpy() for o in ops])
    return T.numpy(), float(c[0]), want_T, float(want_c)


def test_slicing_with_open_indices():
    for T, c, want_T, want_c in _run(_open_slice_case):
        np.testing.assert_allclose(T, want_T, rtol=1e-11)
        assert abs(c - want_c) < 1e-9


def _slab_case(rank, world):
    rng = np.random.default_rng(7)
    # CP reconstruction "ba,ca,da->bcd" (cfg 4): slab over the first open index
    einstr = "ba,ca,da->bcd"
    ops = [rng.uniform(0, 1, (9, 5)), rng.uniform(0, 1, (7, 5)), rng.uniform(0, 1, (6, 5))]
    T, c, (lo, hi) = parallel.contract_output_slabs(einstr, *ops, contract_fn=_oracle_contract)
    T2, c2, (lo2, hi2) = parallel.contract_output_slabs(einstr, *ops, label="c", contract_fn=_oracle_contract)
    want_T, want_c = _oracle_contract(einstr, *ops)
    return T.numpy(), float(c[0]), want_T[lo:hi], T2.numpy(), float(c2[0]), want_T[:, lo2:hi2], float(want_c)


def test_output_slab_sharding_keeps_the_global_scale():
    for T, c, want_T, T2, c2, want_T2, want_c in _run(_slab_case):
        np.testing.assert_allclose(T, want_T, rtol=1e-12)
        np.testing.assert_allclose(T2, want_T2, rtol=1e-12)
        assert abs(c - want_c) < 1e-9 and abs(c2 - want_c) < 1e-9


def test_sliced_network_bookkeeping():
    net = parallel.SlicedNetwork("abc,bd,cde->ae", [(5, 6, 7), (6, 8), (7, 8, 3)], ("d", "b"))
    assert net.einstr == "ac,,ce->ae"
    assert net.n_slices == 48 and net.shapes == ((5, 7), (), (7, 3))
    assert net.slice_index(0) == (0, 0) and net.slice_index(47) == (7, 5)
    with pytest.raises(ValueError):
        parallel.SlicedNetwork("ab,bc->ac", [(2, 3), (3, 4)], ("a",))
    assert parallel.slice_assignments(10, rank=1, world=4) == [1, 5, 9]
    # single process: merges are the identity / a plain re-canonicalisation
    T = torch.tensor([2.0, -6.0])
    out, c = parallel.merge_slice_sums(T, torch.tensor(1.0))
    np.testing.assert_allclose(out.numpy(), [0.5, -1.5])
    assert abs(float(c[0]) - (1.0 + np.log(4.0))) < 1e-12


def test_choose_slice_labels_prefers_big_intermediates():
    einstr = "ab,acd,ce,bfg,dfhi,ehj,gk,ikl,jl->"
    shapes = [(8, 8), (8, 8, 8), (8, 8), (8, 8, 8), (8, 8, 8, 8), (8, 8, 8), (8, 8), (8, 8, 8), (8, 8)]
    lst = planner.contract_path(einstr, *shapes)[1]
    labels = parallel.choose_slice_labels(einstr, shapes, lst, 64)
    assert len(labels) == 2 and len(set(labels)) == 2
    biggest = max((c[2].split("->")[1] for c in lst[:-1]), key=len)
    assert labels[0] in biggest
