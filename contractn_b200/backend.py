"""Seam A: a per-op backend module for the UNMODIFIED reference.

The reference resolves every array operation of its executor by name in a module chosen with
``contract(..., backend=NAME)`` (``contractn/einsum.py:24-28`` -> ``get_func(name, backend)`` =
``getattr(importlib.import_module(backend), name)``).  Passing ``backend="contractn_b200.backend"``
therefore runs each pairwise ``tensordot`` / ``einsum`` of ``_core_contract`` (einsum.py:371, 382) on
the B200 -- one device plan per call, result handed back as a host array -- with zero patching.
One library call per step means no fusion across steps and one PCIe round trip per operation: this
seam is for bring-up and step-by-step parity triage; the performance seam is B
(``contractn_b200.install`` / ``core_contract``).

``tensordot``, ``transpose`` and ``einsum`` carry the names and argument conventions the reference
calls them with (einsum.py:365-384); the stabilisation helpers (einsum.py:32-86: ``zeros``, ``ones``,
``where``, ``abs``, ``log``, ``exp``, ``sum``) act on the host arrays those calls return.
"""
import numpy as np

from . import engine, planner

_LETTERS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def _dense(einstr, *ops):
    """One einsum on the device, un-normalised result: the engine's split pair multiplied back by the
    plan's last kernel (CTN_OUT_DENSE), i.e. the product itself to the rounding of one multiply."""
    ops = [np.asarray(o) for o in ops]
    shapes = tuple(o.shape for o in ops)
    lst = planner.cached_contract_list(einstr, shapes, optimize="auto")
    dtype = engine._compute_dtype({o.dtype for o in ops})
    if dtype == np.float32:
        T, c = engine._run_numpy(ops, lst)
        with np.errstate(over="ignore", invalid="ignore"):
            return (T.astype(np.float64) * np.exp(c)).astype(np.float32)
    return engine._run_numpy(ops, lst, engine.CTN_OUT_DENSE)[0]


def tensordot(x, y, axes=2):
    """``numpy.tensordot`` semantics (the reference passes ``axes=((left...), (right...))``, einsum.py:365)."""
    x, y = np.asarray(x), np.asarray(y)
    if isinstance(axes, int):
        ax, ay = list(range(x.ndim - axes, x.ndim)), list(range(axes))
    else:
        ax, ay = axes
        ax = [ax] if isinstance(ax, int) else list(ax)
        ay = [ay] if isinstance(ay, int) else list(ay)
    ax = [a % x.ndim for a in ax]
    ay = [a % y.ndim for a in ay]
    sx = [None] * x.ndim
    sy = [None] * y.ndim
    nxt = iter(_LETTERS)
    for i, j in zip(ax, ay):
        sx[i] = sy[j] = next(nxt)
    for i in range(x.ndim):
        if sx[i] is None:
            sx[i] = next(nxt)
    for j in range(y.ndim):
        if sy[j] is None:
            sy[j] = next(nxt)
    out = [sx[i] for i in range(x.ndim) if i not in ax] + [sy[j] for j in range(y.ndim) if j not in ay]
    return _dense("".join(sx) + "," + "".join(sy) + "->" + "".join(out), x, y)


def einsum(subscripts, *operands, **kwargs):
    """``numpy.einsum(subscripts, *operands)`` on the device (``dtype`` / ``order`` / ``casting`` accepted)."""
    if kwargs.get("dtype") is not None:
        operands = [np.asarray(o, dtype=kwargs["dtype"]) for o in operands]
    return _dense(subscripts, *operands)


def transpose(x, axes=None):
    return np.transpose(x, axes)  # a view, as in the reference (einsum.py:374-377)


zeros, ones, where, log, exp, allclose = np.zeros, np.ones, np.where, np.log, np.exp, np.allclose
abs = np.abs  # noqa: A001  (the reference looks these names up verbatim, einsum.py:56-86)
sum = np.sum  # noqa: A001
