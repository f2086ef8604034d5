"""Boundary-sharded sweep (contractn_b200.sharded): host logic on CPU.

The segments are contracted by the oracle (the product path needs a GPU); under test are
the segmentation of the path, the label/layout bookkeeping of the all-to-all re-sharding
and the scale merges -- single process (world 1: every exchange is the identity, the
split/merge views are still exercised) and world_size-2 gloo.
"""
import numpy as np
import pytest
import torch

from contractn_b200 import networks, planner, sharded
from oracle import contract_oracle as oracle
from test_parallel_gloo import _oracle_contract, _run


def _grid(L, D, seed=5):
    tn, path = networks.peps_tn(L, D, seed=seed)
    return tn.einsum_str, [np.asarray(p) for p in tn.params], path


def _open_chain(seed=6):
    """A chain with three open legs that survive to the output (ends sharded: 'final' exchange)."""
    rng = np.random.default_rng(seed)
    ops = [rng.uniform(0.1, 1.0, s) for s in [(4, 6), (6, 8, 5), (5, 4, 7), (7, 6, 3), (3, 2)]]
    einstr = "ab,bcd,def,fgh,hi->gceai"
    path = ((0, 1), (0, 3), (0, 2), (0, 1))
    return einstr, ops, path


def _batched_mps(seed=7):
    tn, path = networks.mps_tn(sites=6, bond=4, phys=2, batch=8, dtype=np.float64, seed=seed)
    return tn.einsum_str, [np.asarray(p) for p in tn.params], path


CASES = {"grid4": (lambda: _grid(4, 4), 64), "grid3": (lambda: _grid(3, 2), 8),
         "open_chain": (_open_chain, 32), "batched_mps": (_batched_mps, 16)}


def _check(name, world_expected=None):
    make, min_numel = CASES[name]
    einstr, ops, path = make()
    T, c = sharded.contract_sharded_sweep(einstr, *[torch.as_tensor(o) for o in ops], optimize=path,
                                          min_numel=min_numel, contract_fn=_oracle_contract)
    Tr, cr = _oracle_contract(einstr, *ops, optimize=path)
    Tr = np.asarray(Tr)
    assert tuple(T.shape) == Tr.shape
    np.testing.assert_allclose(T.numpy(), Tr, rtol=1e-12, atol=0)
    assert abs(float(c.reshape(-1)[0]) - float(cr)) < 1e-9
    return True


@pytest.mark.parametrize("name", sorted(CASES))
def test_world1_matches_oracle(name):
    assert _check(name)


def test_segmentation_of_a_grid():
    einstr, ops, path = _grid(4, 4)
    shapes = tuple(o.shape for o in ops)
    lst = planner.cached_contract_list(einstr, shapes, optimize=path)
    segs, n, perm = sharded.plan_segments(einstr, shapes, lst, 2, min_numel=64)
    kinds = [s.exchange for s in segs]
    assert n == 16 and perm is None
    assert kinds.count("a2a") >= 2 and kinds.count("gather") == 1 and kinds[-1] is None
    assert sum(len(s.steps) for s in segs) == len(lst)
    for s in segs:
        if s.exchange == "a2a":     # send layout: (s2_hi, s, s2_lo, rest), s2 lives longest among the candidates
            assert s.out[0] == s.hi and s.out[1] == s.s and s.out[2] == s.s2
            assert s.post_labels == s.out[1:] and s.post_local == s.s2
    # the sharded FLOPs are the unsharded FLOPs: every step appears once (split labels have extent W x e/W)
    terms = einstr.split("->")[0].split(",")
    size = {ch: d for t, shp in zip(terms, shapes) for ch, d in zip(t, shp)}
    total = 0
    for s in segs:
        sz = dict(size)
        if s.hi:
            sz[s.hi], sz[s.s2] = 2, size[s.s2] // 2
        shp = tuple(tuple(sz[ch] for ch in t) for t in s.terms)
        sub = planner.cached_contract_list(s.einstr(), shp, optimize=s.path or "auto")
        total += planner.path_flops(sub, s.einstr(), shp)
    assert total == planner.path_flops(lst, einstr, shapes)


def test_no_divisible_label_stays_replicated():
    """Bond 3 on 2 ranks: no leg can be split evenly, every segment is replicated (no exchange), the
    result is still the reference's."""
    einstr, ops, path = _grid(3, 3)
    shapes = tuple(o.shape for o in ops)
    lst = planner.cached_contract_list(einstr, shapes, optimize=path)
    segs, _, _ = sharded.plan_segments(einstr, shapes, lst, 2, min_numel=1)
    assert all(s.mode == "repl" and s.exchange is None for s in segs)
    segs4, _, _ = sharded.plan_segments(einstr, shapes, lst, 4, min_numel=1)   # more ranks than the bond is wide
    assert all(s.mode == "repl" for s in segs4)


def test_chain_input_is_listed_last():
    """The running tensor is the LAST input of every segment after the first, so the planner's step strings put
    the big operand first (rows of the GEMM) exactly as in the unsharded sweep."""
    einstr, ops, path = _grid(4, 4)
    shapes = tuple(o.shape for o in ops)
    lst = planner.cached_contract_list(einstr, shapes, optimize=path)
    segs, n, _ = sharded.plan_segments(einstr, shapes, lst, 2, min_numel=64)
    for s in segs:
        if s.chain_in is not None:
            assert s.inputs[-1] == s.chain_in
            assert all(p[1] == 0 for p in s.path) and s.path[0][0] == len(s.inputs) - 1


def _gloo_case(rank, world, name):
    return _check(name)


@pytest.mark.parametrize("name", sorted(CASES))
def test_world2_gloo_matches_oracle(name):
    import functools

    assert all(_run(functools.partial(_gloo_case, name=name), world=2))


def test_world4_gloo_grid_matches_oracle():
    """Four ranks, bond 4: every sharded leg is split (4, 1) -- the layout of 8 GPUs on bond 8."""
    import functools

    assert all(_run(functools.partial(_gloo_case, name="grid4"), world=4))
