"""Host-side behaviour of the engine front end that needs no GPU."""
import numpy as np
import pytest

from contractn_b200 import engine, networks, planner


def test_unknown_kwargs_raise_like_the_reference():
    with pytest.raises(TypeError):
        engine.contract("ab,bc->ac", np.ones((2, 3)), np.ones((3, 4)), bogus=1)  # einsum.py:286-290


def test_plan_cache_is_keyed_by_device():
    lst = planner.contract_path("ab,bc->ac", (4, 5), (5, 6))[1]
    engine.clear_plan_cache()
    p0 = engine.get_plan(lst, ((4, 5), (5, 6)), np.float64)
    assert engine.get_plan(lst, ((4, 5), (5, 6)), np.float64) is p0
    assert "gemm" in p0.describe() or "interp" in p0.describe()


def test_sweep_paths_are_valid_paths():
    """f3: the explicit sweep-path generators live in the package and produce complete paths."""
    tn, path = networks.mps_tn(6, 4, 2, 3)
    shapes = tuple(p.shape for p in tn.params)
    lst = planner.contract_path(tn.einsum_str, *shapes, optimize=path)[1]
    assert len(lst) == len(shapes) - 1
    tn, path = networks.peps_tn(3, 2)
    shapes = tuple(p.shape for p in tn.params)
    lst = planner.contract_path(tn.einsum_str, *shapes, optimize=path)[1]
    assert len(lst) == len(shapes) - 1 and lst[-1][2].endswith("->")


def test_tf32_plans_prefer_row_fastest_intermediates():
    """The tcgen05 kernel's accumulator rows are TMEM lanes = threads: fused cfg-3 sites are planned with
    row-index-fastest results (coalesced epilogue stores and A-operand reads, no transposition)."""
    tn, path = networks.mps_tn(6, 64, 2, 256)
    shapes = tuple(p.shape for p in tn.params)
    lst = planner.contract_path(tn.einsum_str, *shapes, optimize=path)[1]
    text = engine.Plan(lst, shapes, np.float32).describe()
    fused = [ln for ln in text.splitlines() if "fuse_pr=2" in ln]
    assert len(fused) == 4, text
    assert all("out_strides=1,256" in ln for ln in fused[:-1]), text  # (the last site feeds a generic step)


@pytest.mark.parametrize("einstr,shapes", [
    ("ab,bc,cd->ad", [(3, 4), (4, 0), (0, 5)]),      # the empty index is summed: zeros of shape (3, 5)
    ("ab,bc,cd->ad", [(3, 4), (4, 5), (5, 0)]),      # the empty index is kept: an empty (3, 0) result
    ("ab,bc,cd,de->ae", [(3, 4), (4, 5), (5, 0), (0, 2)]),
    ("ab,cd,be->acde", [(3, 4), (0, 2), (4, 5)]),
])
def test_empty_operands_walk_matches_the_oracle(einstr, shapes):
    """Zero-extent operands (einsum.py:97-106: norm 0 is below min_norm, the register is left alone): the host walk
    over the pairwise list must give the oracle's zeros AND its log-scale, which only the steps among non-empty
    operands feed.  The device runner is replaced by the oracle here (no GPU): what is tested is the bookkeeping."""
    from oracle import contract_oracle as oracle

    rng = np.random.default_rng(3)
    ops = [rng.uniform(0.5, 1.5, s) for s in shapes]
    lst = planner.contract_path(einstr, *shapes)[1]
    want_T, want_c = oracle.core_contract(ops, lst)
    got_T, got_c = engine._run_with_empty_operands(
        ops, lst, engine.CTN_OUT_SPLIT, lambda o, l, m: oracle.core_contract(o, l), lambda shape: np.zeros(shape))
    assert np.asarray(got_T).shape == np.asarray(want_T).shape
    assert not np.any(np.asarray(got_T))
    assert abs(float(got_c) - float(want_c)) < 1e-12
