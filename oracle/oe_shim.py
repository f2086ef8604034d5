"""Stand-in for the un-installable third-party ``opt_einsum`` -- TEST INFRASTRUCTURE.

The reference imports ``opt_einsum`` at module top level
(``contractn/einsum.py:4-5``, ``contractn/utils.py:4``) and the package is not
available in this image (no wheel, no index).  ``install()`` registers a module
of that name exposing only the eight names the reference touches (SURVEY.md
App. A), so the UNMODIFIED reference can be imported from /root/reference to
generate golden vectors (``oracle/make_golden.py``) and to re-run its own tests.

Arithmetic is numpy's / torch's own; path planning is
``contractn_b200.planner`` (opt_einsum 3.3.0 restatement).  If the real
``opt_einsum`` is importable this shim refuses to shadow it.
"""
import importlib
import sys
import types

from contractn_b200 import planner

_VALID = set("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ,->.")


def _torch_backend():
    import torch

    mod = types.ModuleType("opt_einsum.backends.torch")
    mod.transpose = lambda a, axes: a.permute(*axes)
    mod.einsum = lambda eq, *ops: torch.einsum(eq, ops)
    mod.tensordot = lambda x, y, axes=2: torch.tensordot(x, y, dims=axes)
    return mod


_cache = {}


def get_func(func, backend="numpy"):
    key = (backend, func)
    if key not in _cache:
        if backend == "torch":
            lib = _torch_backend()
        else:
            lib = importlib.import_module(backend)
        _cache[key] = getattr(lib, func)
    return _cache[key]


def has_einsum(backend):
    try:
        get_func("einsum", backend)
        return True
    except AttributeError:
        return False


def has_tensordot(backend):
    try:
        get_func("tensordot", backend)
        return True
    except AttributeError:
        return False


def parse_backend(arrays, backend):
    if backend != "auto":
        return backend
    backend = arrays[0].__class__.__module__.split(".")[0]
    if not has_tensordot(backend):
        return "numpy"
    return backend


def _tensordot(x, y, axes, backend="numpy"):
    return get_func("tensordot", backend)(x, y, axes=axes)


def _transpose(x, axes, backend="numpy"):
    return get_func("transpose", backend)(x, axes)


def _einsum(*operands, **kwargs):
    fn = get_func("einsum", kwargs.pop("backend", "numpy"))
    if not isinstance(operands[0], str):
        return fn(*operands, **kwargs)
    eq, operands = operands[0], operands[1:]
    if not set(eq) <= _VALID:
        symbols = sorted(set(eq) - set(",->"))
        table = {s: planner.get_symbol(i) for i, s in enumerate(symbols)}
        eq = "".join(table.get(ch, ch) for ch in eq)
    return fn(eq, *operands, **kwargs)


def install():
    """Register the stand-in as ``opt_einsum`` (no-op if the real one exists)."""
    if "opt_einsum" in sys.modules:
        return sys.modules["opt_einsum"]
    try:
        import opt_einsum  # noqa: F401

        return sys.modules["opt_einsum"]
    except ImportError:
        pass
    oe = types.ModuleType("opt_einsum")
    oe.__version__ = "0+contractn_b200.shim"
    oe.get_symbol = planner.get_symbol
    oe.contract_path = planner.contract_path
    con = types.ModuleType("opt_einsum.contract")
    con.parse_backend = parse_backend
    con._einsum = _einsum
    con._tensordot = _tensordot
    con._transpose = _transpose
    back = types.ModuleType("opt_einsum.backends")
    back.get_func = get_func
    back.has_einsum = has_einsum
    back.has_tensordot = has_tensordot
    oe.contract = con
    oe.backends = back
    sys.modules["opt_einsum"] = oe
    sys.modules["opt_einsum.contract"] = con
    sys.modules["opt_einsum.backends"] = back
    return oe


def import_reference(path="/root/reference"):
    """Import the unmodified reference package (build container only)."""
    install()
    if path not in sys.path:
        sys.path.insert(0, path)
    import contractn

    return contractn
