"""Python host side of the B200 contraction engine.

Mirrors the reference's contraction front end and executor for the one hot path
(`/root/reference/contractn/einsum.py`):

* :func:`contract`       -- same signature and keyword handling as
  ``contract`` (einsum.py:190-310).
* :func:`core_contract`  -- same signature as ``_core_contract``
  (einsum.py:326-393); this is the seam :func:`install` patches into an
  importable reference package (seam B).
* :class:`Plan`          -- a compiled contraction list living on the device
  (thin ctypes wrapper over ``include/ctn_b200.h``).

All arithmetic runs in ``libctn_b200.so`` (hand-written sm_100a kernels).  There
is NO CPU fallback: if the library or a B200 is missing the calls raise.
"""
from __future__ import annotations

import ctypes
import operator
import os
import threading
from collections import OrderedDict

import numpy as np

from . import planner

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libctn_b200.so")

CTN_F32, CTN_F64 = 0, 1
CTN_OUT_SPLIT, CTN_OUT_DENSE = 0, 1
CTN_PLAN_EXACT_STEPS = 1
CTN_PLAN_NO_INTERPRETER = 2
CTN_PLAN_NO_GEMM = 4
CTN_PLAN_RAW_FINAL = 8
CTN_RESULT_DEGENERATE = 1
LAUNCH_KINDS = ("generic", "chain", "interp", "gemm", "finalize", "materialize", "gemm_reduce")

EXPORTS = (
    "ctn_plan_create", "ctn_plan_destroy", "ctn_plan_workspace_bytes",
    "ctn_plan_output_rank", "ctn_plan_output_extents", "ctn_plan_output_numel",
    "ctn_plan_num_launches", "ctn_plan_flops", "ctn_plan_step_bytes",
    "ctn_plan_launch_kind", "ctn_plan_describe", "ctn_plan_execute", "ctn_plan_execute_timed", "ctn_plan_execute_host",
    "ctn_merge_batch_shards_peer", "ctn_merge_batch_shards", "ctn_reshard_peer", "ctn_plan_final_kappa_offset",
    "ctn_status_string", "ctn_last_cuda_error", "ctn_abi_version", "ctn_device_count",
)


CTN_MAILBOX_WORDS = 512  # include/ctn_b200.h


class EngineError(RuntimeError):
    """A call into libctn_b200 failed (status and CUDA error text attached)."""


class _StepDesc(ctypes.Structure):
    _fields_ = [
        ("n_inputs", ctypes.c_int32),
        ("positions", ctypes.POINTER(ctypes.c_int32)),
        ("input_ranks", ctypes.POINTER(ctypes.c_int32)),
        ("input_labels", ctypes.POINTER(ctypes.c_int32)),
        ("out_rank", ctypes.c_int32),
        ("out_labels", ctypes.POINTER(ctypes.c_int32)),
    ]


class _PlanDesc(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("n_operands", ctypes.c_int32),
        ("operand_ranks", ctypes.POINTER(ctypes.c_int32)),
        ("operand_extents", ctypes.POINTER(ctypes.c_int64)),
        ("operand_strides", ctypes.POINTER(ctypes.c_int64)),
        ("n_steps", ctypes.c_int32),
        ("steps", ctypes.POINTER(_StepDesc)),
        ("flags", ctypes.c_int32),
    ]


_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen ``libctn_b200.so`` (rebuilding it first when the sources changed)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        from .build import build_library

        build_library()  # no-op unless the library is missing or older than its sources (content digest)
        lib = ctypes.CDLL(_LIB_PATH)
        vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
        lib.ctn_plan_create.argtypes = [ctypes.POINTER(_PlanDesc), ctypes.POINTER(vp)]
        lib.ctn_plan_create.restype = ctypes.c_int
        lib.ctn_plan_destroy.argtypes = [vp]
        lib.ctn_plan_destroy.restype = None
        lib.ctn_plan_workspace_bytes.argtypes = [vp]
        lib.ctn_plan_workspace_bytes.restype = ctypes.c_size_t
        lib.ctn_plan_output_rank.argtypes = [vp]
        lib.ctn_plan_output_rank.restype = ctypes.c_int
        lib.ctn_plan_output_extents.argtypes = [vp, ctypes.POINTER(i64)]
        lib.ctn_plan_output_extents.restype = ctypes.c_int
        lib.ctn_plan_output_numel.argtypes = [vp]
        lib.ctn_plan_output_numel.restype = i64
        lib.ctn_plan_num_launches.argtypes = [vp]
        lib.ctn_plan_num_launches.restype = ctypes.c_int
        lib.ctn_plan_flops.argtypes = [vp]
        lib.ctn_plan_flops.restype = dbl
        lib.ctn_plan_step_bytes.argtypes = [vp]
        lib.ctn_plan_step_bytes.restype = dbl
        lib.ctn_plan_launch_kind.argtypes = [vp, ctypes.c_int]
        lib.ctn_plan_launch_kind.restype = ctypes.c_int
        lib.ctn_plan_describe.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t]
        lib.ctn_plan_describe.restype = ctypes.c_int
        lib.ctn_plan_execute.argtypes = [vp, ctypes.POINTER(vp), vp, vp, vp, vp, ctypes.c_int, vp]
        lib.ctn_plan_execute.restype = ctypes.c_int
        lib.ctn_plan_execute_timed.argtypes = [vp, ctypes.POINTER(vp), vp, vp, vp, vp, ctypes.c_int, vp,
                                               ctypes.POINTER(ctypes.c_float)]
        lib.ctn_plan_execute_timed.restype = ctypes.c_int
        lib.ctn_plan_execute_host.argtypes = [vp, ctypes.POINTER(vp), vp, ctypes.POINTER(dbl),
                                              ctypes.POINTER(i32), ctypes.c_int]
        lib.ctn_plan_execute_host.restype = ctypes.c_int
        lib.ctn_merge_batch_shards_peer.argtypes = [vp, i64, ctypes.c_int, vp, ctypes.POINTER(vp), ctypes.c_int,
                                                    ctypes.c_int, ctypes.c_uint64, vp, vp]
        lib.ctn_merge_batch_shards_peer.restype = ctypes.c_int
        lib.ctn_reshard_peer.argtypes = [vp, i64, ctypes.c_int, vp, i64, ctypes.POINTER(vp), ctypes.POINTER(vp),
                                         ctypes.POINTER(vp), vp, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, vp, vp]
        lib.ctn_reshard_peer.restype = ctypes.c_int
        lib.ctn_plan_final_kappa_offset.argtypes = [vp]
        lib.ctn_plan_final_kappa_offset.restype = i64
        lib.ctn_merge_batch_shards.argtypes = [vp, i64, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, vp, vp]
        lib.ctn_merge_batch_shards.restype = ctypes.c_int
        lib.ctn_status_string.argtypes = [ctypes.c_int]
        lib.ctn_status_string.restype = ctypes.c_char_p
        lib.ctn_last_cuda_error.argtypes = []
        lib.ctn_last_cuda_error.restype = ctypes.c_char_p
        lib.ctn_abi_version.restype = ctypes.c_int
        lib.ctn_device_count.restype = ctypes.c_int
        _lib = lib
        return lib


def _check(status, what):
    if status != 0:
        lib = load_library()
        msg = lib.ctn_status_string(status).decode()
        cuda = lib.ctn_last_cuda_error().decode()
        raise EngineError(f"{what}: {msg}" + (f" [{cuda}]" if cuda and status == 4 else ""))


def _i32(seq):
    return (ctypes.c_int32 * max(1, len(seq)))(*seq)


def _i64(seq):
    return (ctypes.c_int64 * max(1, len(seq)))(*seq)


class Plan:
    """A contraction list compiled for the device (``ctn_plan`` handle).

    ``contraction_list`` is what ``contract_path(..., einsum_call=True)`` returns
    (einsum.py:341-342); ``shapes``/``strides`` (elements, ``None`` for
    C-contiguous) describe the operands; ``dtype`` is ``np.float32``/``np.float64``.
    """

    def __init__(self, contraction_list, shapes, dtype, strides=None, flags=0):
        lib = load_library()
        self._lib = lib
        self._lock = threading.Lock()  # one workspace / staging set per plan: executions are serialised
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise NotImplementedError(f"dtype {self.dtype} is not supported by the B200 engine")
        self.shapes = tuple(tuple(int(d) for d in s) for s in shapes)
        keep = []  # ctypes arrays must outlive ctn_plan_create
        steps = (_StepDesc * len(contraction_list))()
        for t, (positions, _removed, einsum_str, _rem, _flag) in enumerate(contraction_list):
            ins, out = einsum_str.split("->")
            terms = ins.split(",")
            if len(terms) != len(positions):
                raise ValueError(f"step {t}: {len(positions)} positions for '{einsum_str}'")
            pos = _i32(list(positions))
            ranks = _i32([len(x) for x in terms])
            labels = _i32([ord(ch) for term in terms for ch in term])
            outl = _i32([ord(ch) for ch in out])
            keep += [pos, ranks, labels, outl]
            steps[t].n_inputs = len(terms)
            steps[t].positions = pos
            steps[t].input_ranks = ranks
            steps[t].input_labels = labels
            steps[t].out_rank = len(out)
            steps[t].out_labels = outl
        desc = _PlanDesc()
        desc.abi_version = 1
        desc.dtype = CTN_F32 if self.dtype == np.float32 else CTN_F64
        desc.n_operands = len(self.shapes)
        ranks = _i32([len(s) for s in self.shapes])
        extents = _i64([d for s in self.shapes for d in s])
        desc.operand_ranks = ranks
        desc.operand_extents = extents
        if strides is not None:
            strd = _i64([int(v) for s in strides for v in s])
            desc.operand_strides = strd
            keep.append(strd)
        desc.n_steps = len(contraction_list)
        desc.steps = steps
        desc.flags = flags
        handle = ctypes.c_void_p()
        _check(lib.ctn_plan_create(ctypes.byref(desc), ctypes.byref(handle)), "ctn_plan_create")
        self._h = handle
        self.flags = flags
        rank = lib.ctn_plan_output_rank(handle)
        ext = (ctypes.c_int64 * max(1, rank))()
        _check(lib.ctn_plan_output_extents(handle, ext), "ctn_plan_output_extents")
        self.out_shape = tuple(int(ext[i]) for i in range(rank))
        self.workspace_bytes = int(lib.ctn_plan_workspace_bytes(handle))
        self.final_kappa_offset = int(lib.ctn_plan_final_kappa_offset(handle))  # -1: output is canonical
        self.num_launches = int(lib.ctn_plan_num_launches(handle))
        self.flops = float(lib.ctn_plan_flops(handle))
        self.step_bytes = float(lib.ctn_plan_step_bytes(handle))
        kinds = []
        i = 0
        while True:
            k = lib.ctn_plan_launch_kind(handle, i)
            if k < 0:
                break
            kinds.append(LAUNCH_KINDS[k])
            i += 1
        self.launch_kinds = tuple(kinds)
        self._dev = None  # torch scratch for the device path

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and self._lib is not None:
            try:
                self._lib.ctn_plan_destroy(h)
            except Exception:
                pass

    # -- host buffers in, host buffers out ---------------------------------
    def execute_host(self, arrays, out_mode=CTN_OUT_SPLIT, ptrs=None):
        """Run on host ``numpy`` arrays (copies in and out).  Returns
        ``(result, log_scale, flags)``."""
        n = len(self.shapes)
        assert len(arrays) == n
        if ptrs is None:
            ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrays])
        out = np.empty(self.out_shape, dtype=self.dtype)
        log_scale = ctypes.c_double(0.0)
        flags = ctypes.c_int32(0)
        with self._lock:
            _check(
                self._lib.ctn_plan_execute_host(self._h, ptrs, out.ctypes.data, ctypes.byref(log_scale),
                                                ctypes.byref(flags), out_mode),
                "ctn_plan_execute_host",
            )
        return out, log_scale.value, flags.value

    # -- device buffers (torch CUDA tensors or raw pointers) ----------------
    def _device_scratch(self, device):
        import torch

        if self._dev is None or self._dev[0] != device:
            ws = torch.empty(self.workspace_bytes + 256, dtype=torch.uint8, device=device)
            base = ws.data_ptr()
            pad = (-base) % 256
            tail = torch.zeros(2, dtype=torch.float64, device=device)  # [log_scale, flags bits]
            self._dev = (device, ws, base + pad, tail)
        return self._dev

    def execute_device(self, tensors, out=None, out_mode=CTN_OUT_SPLIT, stream=None, timed=False):
        """Run on torch CUDA tensors without leaving the device.  Returns
        ``(out, tail)`` where ``tail[0]`` is the log-scale (float64, on device) and
        ``tail[1]`` reinterpreted as int32 holds the result flags.  With
        ``timed=True`` returns ``(out, tail, [ms per kernel])`` (synchronises)."""
        import torch

        device = tensors[0].device
        # the plan's descriptor tables live on the device that was current when it was created
        # (get_plan keys on it); launches go to that device whatever the caller's current one is.
        # (torch.cuda.device() as a context manager costs ~0.1 ms per call: only switch when needed)
        prev = torch.cuda.current_device()
        switch = device.index is not None and prev != device.index
        if switch:
            torch.cuda.set_device(device)
        try:
            with self._lock:
                return self._execute_device_locked(tensors, device, out, out_mode, stream, timed)
        finally:
            if switch:
                torch.cuda.set_device(prev)

    def _execute_device_locked(self, tensors, device, out, out_mode, stream, timed):
        import torch

        _, _ws, ws_ptr, tail = self._device_scratch(device)
        tdtype = torch.float32 if self.dtype == np.float32 else torch.float64
        if out is None:
            out = torch.empty(self.out_shape, dtype=tdtype, device=device)
        n = len(self.shapes)
        ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in tensors])
        if stream is None:
            stream = torch.cuda.current_stream(device).cuda_stream
        if timed:
            ms = (ctypes.c_float * max(1, self.num_launches))()
            _check(
                self._lib.ctn_plan_execute_timed(self._h, ptrs, out.data_ptr(), tail.data_ptr(),
                                                 tail.data_ptr() + 8, ws_ptr, out_mode, stream, ms),
                "ctn_plan_execute_timed",
            )
            return out, tail, [float(ms[i]) for i in range(self.num_launches)]
        _check(
            self._lib.ctn_plan_execute(self._h, ptrs, out.data_ptr(), tail.data_ptr(), tail.data_ptr() + 8,
                                       ws_ptr, out_mode, stream),
            "ctn_plan_execute",
        )
        return out, tail

    def bind(self, tensors, out=None, out_mode=CTN_OUT_SPLIT):
        """Pre-resolve everything a device-resident call needs (operand pointers, output, workspace) so
        that repeated executions on the SAME resident operands cost one ctypes call each -- collecting 100
        ``data_ptr()`` values per call is otherwise as expensive as the 0.3 ms cfg-2 step itself.  The
        operands are kept alive by the returned object; their contents may change between runs."""
        import torch

        device = tensors[0].device
        with torch.cuda.device(device):
            _, _ws, ws_ptr, tail = self._device_scratch(device)
            tdtype = torch.float32 if self.dtype == np.float32 else torch.float64
            if out is None:
                out = torch.empty(self.out_shape, dtype=tdtype, device=device)
        return BoundCall(self, list(tensors), out, tail, ws_ptr, out_mode, device)

    def describe(self):
        """What the plan compiler decided: one line per launch."""
        buf = ctypes.create_string_buffer(1 << 20)
        _check(self._lib.ctn_plan_describe(self._h, buf, len(buf)), "ctn_plan_describe")
        return buf.value.decode()

    @property
    def may_degenerate(self):
        # fused chains (any dtype) and fused GEMM+consumer steps (float32) can flag a degenerate norm
        fused = "chain" in self.launch_kinds or (self.dtype == np.float32 and "gemm" in self.launch_kinds)
        return fused and not (self.flags & CTN_PLAN_EXACT_STEPS)


class BoundCall:
    """A plan bound to resident operands (see :meth:`Plan.bind`).  ``run()`` enqueues one execution on
    torch's current stream and returns ``(out, tail)`` -- the same buffers every time."""

    def __init__(self, plan, tensors, out, tail, ws_ptr, out_mode, device):
        self.plan, self.tensors, self.out, self.tail = plan, tensors, out, tail
        self._ptrs = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        self._args = (plan._h, self._ptrs, out.data_ptr(), tail.data_ptr(), tail.data_ptr() + 8, ws_ptr, out_mode)
        self._device = device
        self._fn = plan._lib.ctn_plan_execute

    def run(self, stream=None):
        import torch

        if stream is None:
            stream = torch.cuda.current_stream(self._device).cuda_stream
        with self.plan._lock:
            _check(self._fn(*self._args, stream), "ctn_plan_execute")
        return self.out, self.tail


# ---------------------------------------------------------------------------
# plan cache (plays the role of the lru_cache on _contract_path, einsum.py:313)
# ---------------------------------------------------------------------------
_plan_cache = OrderedDict()
_PLAN_CACHE_SIZE = 128


def _hashable_list(contraction_list):
    return tuple((tuple(c[0]), c[2]) for c in contraction_list)


_plan_cache_lock = threading.Lock()


def get_plan(contraction_list, shapes, dtype, strides=None, flags=0, device=None):
    """Cached :class:`Plan`.  ``device`` is the CUDA device index the plan is bound to (its
    descriptor tables are allocated there); ``None`` = the current device (host-buffer path)."""
    key = (_hashable_list(contraction_list), tuple(shapes), np.dtype(dtype).str,
           None if strides is None else tuple(strides), flags, device)
    with _plan_cache_lock:
        plan = _plan_cache.get(key)
        if plan is not None:
            _plan_cache.move_to_end(key)
            return plan
    if device is None:
        plan = Plan(contraction_list, shapes, dtype, strides=strides, flags=flags)
    else:
        import torch

        with torch.cuda.device(device):
            plan = Plan(contraction_list, shapes, dtype, strides=strides, flags=flags)
    with _plan_cache_lock:
        plan = _plan_cache.setdefault(key, plan)
        if len(_plan_cache) > _PLAN_CACHE_SIZE:
            _plan_cache.popitem(last=False)
    return plan


def clear_plan_cache():
    with _plan_cache_lock:
        _plan_cache.clear()


# ---------------------------------------------------------------------------
# operand handling
# ---------------------------------------------------------------------------
def _is_torch(x):
    return type(x).__module__.split(".")[0] == "torch"


def _compute_dtype(dtypes):
    kinds = {np.dtype(d).kind for d in dtypes}
    if "c" in kinds:
        raise NotImplementedError("complex tensors are not supported by the B200 engine")
    res = np.result_type(*dtypes) if dtypes else np.dtype(np.float64)
    if res == np.float32 or res == np.float16:
        return np.dtype(np.float32)
    return np.dtype(np.float64)


def _prep_numpy(op, dtype):
    a = np.asarray(op)
    if a.dtype != dtype:
        a = a.astype(dtype)
    item = a.dtype.itemsize
    if any(s < 0 or s % item for s in a.strides):
        a = np.ascontiguousarray(a)
    return a


def _elem_strides(a):
    item = a.dtype.itemsize
    return tuple(int(s // item) for s in a.strides)


# Networks of many small tensors (cfg 1: 1001 operands) spend milliseconds per call on per-operand
# Python work (dtype, strides, ctypes pointers).  When a call passes the SAME array objects as the
# previous one -- what ``TN.contract()`` / ``contract_fun(tn.params, ...)`` do -- that work is reused:
# the bundle keeps the arrays alive (so the pointers stay valid) and is re-validated by identity,
# shape and strides.  Only unconverted operands are cached (values are always read from the caller's buffers).
_BUNDLE_MIN_OPERANDS = 16
_BUNDLE_MAX_BYTES = 16 << 20
_bundle = None


def _run_with_empty_operands(operands, contraction_list, out_mode, runner, zeros):
    """Operands with a zero extent (einsum.py:97-106: ``sum(abs(T))`` of an empty or all-zero tensor is 0, below
    ``min_norm``, so such a step leaves the register alone).  The plan compiler has no kernels for empty tensors, so
    the pairwise list is walked on the host: a step that touches an empty operand -- or the zeros that came out of
    such a step -- yields zeros of its output shape and no scale; every other step is ONE device contraction through
    ``runner`` (its operands are raw inputs or canonical results, so its log-scale is exactly what the reference's
    ``stabilize`` would add at that step).  Only the list's bookkeeping happens here: no arithmetic on data."""
    running = [(op, 0 in tuple(op.shape)) for op in operands]
    log_scale = 0.0
    for step in contraction_list:
        inds, einsum_str = step[0], step[2]
        popped = [running.pop(x) for x in sorted(inds, reverse=True)]
        ins, out = einsum_str.split("->")
        ins = ins.split(",")
        if any(dead for _, dead in popped):
            dims = {}
            for (arr, _), lab in zip(popped, ins):
                dims.update(zip(lab, tuple(arr.shape)))
            running.append((zeros(tuple(dims[ch] for ch in out)), True))
        else:
            # the local one-step list pops its operands in descending position order (einsum.py:347): hand them
            # over reversed so that they meet the step's einsum string in order
            local = [(tuple(range(len(popped) - 1, -1, -1)), step[1], einsum_str, (out,), step[4])]
            t, c = runner([arr for arr, _ in popped][::-1], local, CTN_OUT_SPLIT)
            log_scale += float(c)
            running.append((t, False))
    (result, dead), = running
    if not dead and out_mode == CTN_OUT_DENSE:  # (cannot happen with an empty operand in the list; kept for symmetry)
        result = result * float(np.exp(log_scale))
    return result, log_scale


def _run_complex_numpy(arrays, contraction_list, out_mode):
    import math

    from . import complex_lowering as cl

    arrays = [np.asarray(a) for a in arrays]
    single = np.result_type(*[a.dtype for a in arrays]) == np.complex64
    rdt, cdt = (np.float32, np.complex64) if single else (np.float64, np.complex128)

    def to_real(x):
        x = np.ascontiguousarray(x, dtype=cdt)
        return x.view(rdt).reshape(x.shape + (2,))

    def to_complex(x):
        x = np.ascontiguousarray(x, dtype=rdt)
        return x.view(cdt).reshape(x.shape[:-1])

    def modulus_mean(x):
        tot = float(np.abs(x).sum(dtype=np.float64))
        return tot, tot / max(x.size, 1)

    out, log_scale = cl.run(arrays, contraction_list, lambda e, *o, **k: contract(e, *o, split_format=True, **k),
                            to_real, cl.structure_tensor(rdt), to_complex, modulus_mean, math.log)
    if out_mode == CTN_OUT_DENSE:
        with np.errstate(over="ignore", invalid="ignore"):
            out = out * np.exp(log_scale)
    return out.astype(cdt, copy=False), np.array(log_scale, dtype=np.float64)


def _run_numpy(arrays, contraction_list, out_mode=CTN_OUT_SPLIT):
    global _bundle
    if any(np.iscomplexobj(a) for a in arrays):
        if any(0 in np.shape(a) for a in arrays):
            raise NotImplementedError("complex operands with a zero extent")
        return _run_complex_numpy(arrays, contraction_list, out_mode)
    if any(0 in np.shape(a) for a in arrays):
        dtype = _compute_dtype({np.asarray(a).dtype for a in arrays})
        out, log_scale = _run_with_empty_operands([np.asarray(a) for a in arrays], contraction_list, out_mode,
                                                  _run_numpy, lambda shape: np.zeros(shape, dtype=dtype))
        return np.asarray(out, dtype=dtype), np.array(log_scale, dtype=np.float64)
    b = _bundle
    if (b is not None and b[0] is contraction_list and len(arrays) == len(b[1])
            and all(map(operator.is_, arrays, b[1]))
            and tuple([a.shape for a in arrays]) == b[2] and tuple([a.strides for a in arrays]) == b[3]):
        arrays, shapes, plan, ptrs, dtype, strides = b[1], b[2], b[4], b[5], b[6], b[7]
    else:
        src = arrays
        dtype = _compute_dtype({np.asarray(a).dtype for a in arrays})
        arrays = [_prep_numpy(a, dtype) for a in arrays]
        shapes = tuple([a.shape for a in arrays])
        strides = tuple([_elem_strides(a) for a in arrays])
        plan = get_plan(contraction_list, shapes, dtype, strides=strides)
        ptrs = (ctypes.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
        _bundle = None
        if (len(arrays) >= _BUNDLE_MIN_OPERANDS and all(map(operator.is_, arrays, src))
                and sum(a.nbytes for a in arrays) <= _BUNDLE_MAX_BYTES):
            _bundle = (contraction_list, list(arrays), shapes, tuple([a.strides for a in arrays]), plan, ptrs,
                       dtype, strides)
    out, log_scale, flags = plan.execute_host(arrays, out_mode=out_mode, ptrs=ptrs)
    if flags & CTN_RESULT_DEGENERATE:
        exact = get_plan(contraction_list, shapes, dtype, strides=strides, flags=CTN_PLAN_EXACT_STEPS)
        out, log_scale, flags = exact.execute_host(arrays, out_mode=out_mode, ptrs=ptrs)
    return out, np.array(log_scale, dtype=np.float64)


_AUTOGRAD_FN = None


def _contract_split_nograd(einstr, *operands):
    import torch

    with torch.no_grad():
        return contract(einstr, *operands, split_format=True)


def _run_torch(tensors, contraction_list, out_mode=CTN_OUT_SPLIT):
    """torch tensors in, torch tensors out.  Operands that require gradients go through
    ``autograd.ContractFn`` (the reference's torch backend is differentiable plain torch ops,
    einsum.py:326-393): the backward is one more stabilised contraction per such operand."""
    import torch

    if torch.is_grad_enabled() and any(t.requires_grad for t in tensors):
        if any(t.is_complex() for t in tensors):
            raise NotImplementedError("autograd through a contraction of complex tensors")
        global _AUTOGRAD_FN
        if _AUTOGRAD_FN is None:
            from . import autograd

            _AUTOGRAD_FN = autograd.make_function(_run_torch_nograd, _contract_split_nograd)
        return _AUTOGRAD_FN.apply(contraction_list, out_mode, out_mode == CTN_OUT_DENSE, *tensors)
    return _run_torch_nograd(tensors, contraction_list, out_mode)


def _run_complex_torch(tensors, contraction_list, out_mode):
    import math

    import torch

    from . import complex_lowering as cl

    single = all(t.dtype in (torch.complex64, torch.float32, torch.float16) for t in tensors)
    rdt, cdt = (torch.float32, torch.complex64) if single else (torch.float64, torch.complex128)
    dev = tensors[0].device

    def modulus_mean(x):
        tot = float(x.abs().sum(dtype=torch.float64))
        return tot, tot / max(x.numel(), 1)

    with torch.no_grad():
        out, log_scale = cl.run(
            [t.detach() for t in tensors], contraction_list,
            lambda e, *o, **k: contract(e, *o, split_format=True, **k),
            lambda x: torch.view_as_real(x.to(cdt).contiguous()),
            torch.as_tensor(cl.structure_tensor(np.float32 if single else np.float64), device=dev),
            lambda x: torch.view_as_complex(x.to(rdt).contiguous()), modulus_mean, math.log)
        if out_mode == CTN_OUT_DENSE:
            out = out * torch.exp(torch.tensor(log_scale, dtype=torch.float64, device=dev)).to(rdt)
    return out, torch.tensor(log_scale, dtype=rdt, device=dev)


def _run_torch_nograd(tensors, contraction_list, out_mode=CTN_OUT_SPLIT):
    import torch

    if not tensors[0].is_cuda:
        out, log_scale = _run_numpy([t.detach().numpy() for t in tensors], contraction_list, out_mode)
        res = torch.from_numpy(np.ascontiguousarray(out))
        return res, torch.tensor(float(log_scale), dtype=torch.float32 if res.dtype in (torch.float32, torch.complex64)
                                 else torch.float64)
    tdt = torch.float64 if any(t.dtype == torch.float64 for t in tensors) else torch.float32
    if any(t.is_complex() for t in tensors):
        return _run_complex_torch(tensors, contraction_list, out_mode)
    tensors = [t.detach() for t in tensors]
    if any(0 in tuple(t.shape) for t in tensors):
        out, log_scale = _run_with_empty_operands(
            tensors, contraction_list, out_mode, _run_torch_nograd,
            lambda shape: torch.zeros(shape, dtype=tdt, device=tensors[0].device))
        return out.to(tdt), torch.tensor(log_scale, dtype=tdt, device=tensors[0].device)
    tensors = [t if t.dtype == tdt else t.to(tdt) for t in tensors]
    tensors = [t if all(s >= 0 for s in t.stride()) else t.contiguous() for t in tensors]
    shapes = tuple(tuple(t.shape) for t in tensors)
    strides = tuple(tuple(t.stride()) for t in tensors)
    ndt = np.float64 if tdt == torch.float64 else np.float32
    dev = tensors[0].device.index
    plan = get_plan(contraction_list, shapes, ndt, strides=strides, device=dev)
    out, tail = plan.execute_device(tensors, out_mode=out_mode)
    if plan.may_degenerate:
        # one 4-byte read back: only plans with fused steps can raise the flag (see Plan.may_degenerate)
        flags = int(tail[1:2].view(torch.int32)[0].item())
        if flags & CTN_RESULT_DEGENERATE:
            exact = get_plan(contraction_list, shapes, ndt, strides=strides, flags=CTN_PLAN_EXACT_STEPS, device=dev)
            out, tail = exact.execute_device(tensors, out_mode=out_mode)
    # (a copy: `tail` is the plan's own output slot and is overwritten by its next execution)
    return out, tail[0].to(tdt, copy=True)


def core_contract(operands, contract_list, backend=None, **einsum_kwargs):
    """Drop-in for ``_core_contract`` (einsum.py:326-393): ``(tensor, log_scale)``.

    ``operands`` are borrowed and never modified.  ``backend`` is accepted for
    signature compatibility; the result lives where the operands live (numpy in ->
    numpy out, torch CUDA in -> torch CUDA out).
    """
    operands = list(operands)
    if einsum_kwargs.get("dtype") is not None:
        operands = [np.asarray(o, dtype=einsum_kwargs["dtype"]) for o in operands]
    if _is_torch(operands[0]):
        return _run_torch(operands, contract_list)
    return _run_numpy(operands, contract_list)


def contract(*operands, **kwargs):
    """Stabilised einsum on the B200: mirror of ``contract`` (einsum.py:190-310).

    ``contract(einstr, *operands, optimize="auto", split_format=False,
    memory_limit=None, use_blas=True, backend="auto", dtype=, order=, casting=)``
    """
    optimize_arg = kwargs.pop("optimize", True)
    if optimize_arg is True:
        optimize_arg = "auto"
    valid_einsum_kwargs = ("dtype", "order", "casting")
    einsum_kwargs = {k: v for k, v in kwargs.items() if k in valid_einsum_kwargs}
    use_blas = kwargs.pop("use_blas", True)
    split_format = kwargs.pop("split_format", False)
    memory_limit = kwargs.pop("memory_limit", None)
    kwargs.pop("backend", "auto")
    unknown = [k for k in kwargs if k not in valid_einsum_kwargs]
    if unknown:
        raise TypeError("Did not understand the following kwargs: {}".format(unknown))

    assert isinstance(operands[0], str)
    einstr, ops = operands[0], operands[1:]
    shapes = tuple(tuple(op.shape) for op in ops)
    contract_list = planner.cached_contract_list(
        einstr, shapes, optimize=optimize_arg, memory_limit=memory_limit, use_blas=use_blas
    )
    if split_format:
        return core_contract(ops, contract_list, **einsum_kwargs)
    # destabilize (einsum.py:110-114, 310): T-hat * exp(c), applied by the plan's last kernel
    # (CTN_OUT_DENSE) -- no extra pass, no host arithmetic
    ops = list(ops)
    if einsum_kwargs.get("dtype") is not None:
        ops = [np.asarray(o, dtype=einsum_kwargs["dtype"]) for o in ops]
    if _is_torch(ops[0]):
        return _run_torch(ops, contract_list, CTN_OUT_DENSE)[0]
    if any(np.iscomplexobj(a) for a in ops):
        return _run_numpy(ops, contract_list, CTN_OUT_DENSE)[0]
    dtype = _compute_dtype({np.asarray(a).dtype for a in ops})
    if dtype == np.float32:
        # numpy's own rule for float32 array * 0-d float64 register (einsum.py:114, 338) is a float64
        # result: keep that dtype by multiplying the float32 mantissas on the host
        result, log_scale = _run_numpy(ops, contract_list)
        with np.errstate(over="ignore", invalid="ignore"):
            return result * np.exp(log_scale)
    return _run_numpy(ops, contract_list, CTN_OUT_DENSE)[0]


def install(contractn_module=None):
    """Patch an importable reference package so that ``TN.contract`` runs on the
    B200 (seam B: replaces ``contractn.einsum._core_contract``).  Returns the
    original function so the caller can restore it."""
    if contractn_module is None:
        import contractn as contractn_module
    ein = contractn_module.einsum
    original = ein._core_contract
    ein._core_contract = core_contract
    return original


def device_available():
    """True when the library loads and sees at least one compute-10.x device."""
    try:
        return load_library().ctn_device_count() > 0
    except Exception:
        return False
