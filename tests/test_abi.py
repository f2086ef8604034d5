"""The C-ABI library: builds, loads, exports what include/ctn_b200.h declares,
and compiles plans without a GPU (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from contractn_b200 import engine, planner
from contractn_b200.build import build_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ctn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ctn_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in ctn_b200.h but not exported"
    assert set(names) == set(engine.EXPORTS)
    assert engine.load_library().ctn_abi_version() == 1


def test_status_strings():
    lib = engine.load_library()
    assert lib.ctn_status_string(0) == b"ok"
    for code in range(1, 7):
        assert lib.ctn_status_string(code) not in (b"ok", b"unknown status")


def _plan(einstr, shapes, dtype=np.float64, flags=0, **kw):
    lst = planner.contract_path(einstr, *shapes, **kw)[1]
    return engine.Plan(lst, shapes, dtype, flags=flags)


def test_plan_introspection_hub_is_one_fused_chain():
    p = _plan(",".join(["ab"] * 100) + "->ab", [(2, 4096)] * 100)
    assert p.out_shape == (2, 4096)
    assert p.launch_kinds == ("chain", "materialize")
    assert p.num_launches == 2
    assert p.flops == 99 * 2 * 4096
    assert p.step_bytes == 101 * 2 * 4096 * 8  # every operand once + the output once
    exact = _plan(",".join(["ab"] * 100) + "->ab", [(2, 4096)] * 100, flags=engine.CTN_PLAN_EXACT_STEPS)
    assert exact.launch_kinds == ("chain",) * 99 + ("finalize",)


def test_plan_introspection_tiny_chain_is_one_launch():
    n = 1000
    terms = ["a"] + [planner.get_symbol(i) + planner.get_symbol(i + 1) for i in range(n)]
    p = _plan(",".join(terms) + "->" + planner.get_symbol(n), [(3,)] + [(3, 3)] * n)
    assert p.launch_kinds == ("interp",)
    assert p.num_launches == 1
    assert p.out_shape == (3,)


def test_plan_flops_follow_opt_einsum_convention():
    shapes = [(30, 40), (40, 50), (50, 60)]
    lst = planner.contract_path("ab,bc,cd->ad", *shapes)[1]
    p = engine.Plan(lst, shapes, np.float64)
    assert p.flops == planner.path_flops(lst, "ab,bc,cd->ad", shapes)


def test_plan_errors_are_status_codes_not_crashes():
    lst = planner.contract_path("ab,bc->ac", (3, 4), (4, 5))[1]
    with pytest.raises(engine.EngineError, match="disagree"):
        engine.Plan(lst, [(3, 4), (7, 5)], np.float64)
    with pytest.raises(engine.EngineError, match="invalid"):
        engine.Plan(lst, [(3, 4, 2), (4, 5)], np.float64)
    with pytest.raises(NotImplementedError):
        engine.Plan(lst, [(3, 4), (4, 5)], np.complex128)
    bad = [((5, 0), set(), "ab,bc->ac", None, "GEMM")]
    with pytest.raises(engine.EngineError, match="invalid"):
        engine.Plan(bad, [(3, 4), (4, 5)], np.float64)
    with pytest.raises(engine.EngineError, match="not supported"):
        engine.Plan(lst, [(0, 4), (4, 5)], np.float64)  # empty result


def test_workspace_reuse_keeps_arena_small():
    """Lifetime-based placement: a 64-step chain of equal-size intermediates
    needs two buffers, not 64."""
    n = 64
    terms = [planner.get_symbol(i) + planner.get_symbol(i + 1) for i in range(n)]
    einstr = ",".join(terms) + "->" + planner.get_symbol(0) + planner.get_symbol(n)
    path = tuple([(0, 1)] + [(0, n - 1 - k) for k in range(1, n - 1)])
    shapes = [(128, 128)] * n
    lst = planner.contract_path(einstr, *shapes, optimize=path)[1]
    p = engine.Plan(lst, shapes, np.float64, flags=engine.CTN_PLAN_NO_INTERPRETER)
    assert p.workspace_bytes < 4 * 128 * 128 * 8 + (1 << 20)


def test_header_constants_match_the_host_side():
    import re

    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "ctn_b200.h")).read()
    for name in ("CTN_PLAN_EXACT_STEPS", "CTN_PLAN_NO_INTERPRETER", "CTN_PLAN_NO_GEMM", "CTN_PLAN_RAW_FINAL"):
        m = re.search(name + r"\s*=\s*(\d+)", hdr)
        assert m and int(m.group(1)) == getattr(engine, name), name
    m = re.search(r"#define\s+CTN_MAILBOX_WORDS\s+(\d+)", hdr)
    assert m and int(m.group(1)) == engine.CTN_MAILBOX_WORDS


def test_raw_final_plan_has_no_finalize_launch_and_reports_its_kappa_slot():
    """CTN_PLAN_RAW_FINAL (the plans that feed ctn_reshard_peer): same launches minus the finalize pass, and
    the workspace offset of the pending scale is inside the state block; without the flag the offset is -1."""
    einstr, shapes = "abcde,dexy->abcxy", [(16, 16, 16, 8, 8), (8, 8, 8, 8)]
    lst = planner.contract_path(einstr, *shapes)[1]
    plain = engine.Plan(lst, shapes, np.float64)
    raw = engine.Plan(lst, shapes, np.float64, flags=engine.CTN_PLAN_RAW_FINAL)
    assert plain.launch_kinds == ("gemm", "finalize") and raw.launch_kinds == ("gemm",)
    assert plain.final_kappa_offset == -1
    assert 0 < raw.final_kappa_offset < raw.workspace_bytes and raw.final_kappa_offset % 8 == 0
    # a plan whose last launch finalises by itself ignores the flag
    chain = engine.Plan(planner.contract_path("a,ab,bc->c", (3,), (3, 3), (3, 3))[1], [(3,), (3, 3), (3, 3)],
                        np.float64, flags=engine.CTN_PLAN_RAW_FINAL)
    assert chain.launch_kinds == ("interp",) and chain.final_kappa_offset == -1
