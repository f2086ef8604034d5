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
