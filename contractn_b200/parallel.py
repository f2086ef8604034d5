"""Multi-GPU partitioning of the contraction path (one process per GPU).

The reference has no parallelism at all; these are the two places where the
path shards naturally:

* **batch sharding** -- independent network evaluations along a batch hub are
  split across ranks; each rank contracts its shard to a split-format pair
  ``(T_r, c_r)`` with its OWN scale, and :func:`merge_batch_shards` restores the
  reference's single global scale (the output stays sharded).  On CUDA the whole
  merge is ONE kernel, ``ctn_merge_batch_shards_peer``: the ranks' 24-byte
  triples travel through peer-mapped mailboxes over NVLink, no NCCL call.
* **bond slicing** -- the largest contracted bonds are fixed to one value per
  slice, slices are distributed round-robin, every rank accumulates its slices
  with a log-sum-exp merge and :func:`merge_slice_sums` all-reduces the
  (mantissa, log-scale) pairs over NCCL/NVLink.

* **boundary-sharded sweep** (``contractn_b200.sharded``) -- the bond-parallel path for SPEED: the running
  intermediate is sharded over one of its legs and re-sharded by ONE kernel over peer memory
  (:func:`reshard_peer`); total work stays 1x, where bond slicing multiplies it (it buys memory).

Both merges are exact by the gauge property of the split format (SURVEY.md
finding 0.4): ``T * exp(c)`` is what is summed / concatenated, the final pair is
re-canonicalised to ``mean|T| == 1``.  Works on CUDA tensors over NCCL and on
CPU tensors over gloo (the latter is how the logic is tested without GPUs).
"""
from __future__ import annotations

import math
import os

import torch
import torch.distributed as dist


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


_PEER = {}       # (device index, world, group id) -> peer mailbox state, or False when unavailable
_TABLES = {}


def _peer_state(dev, world, group):
    """Peer-mapped mailboxes for the one-kernel merge: one CTN_MAILBOX_WORDS buffer per rank in symmetric
    memory (``torch.distributed._symmetric_memory``: cuMem handles exchanged once, then plain device
    pointers over NVLink).  Returns None when symmetric memory cannot be set up (then NCCL is used)."""
    key = (dev.index, world, id(group))
    st = _PEER.get(key)
    if st is None:
        st = False
        if os.environ.get("CTN_NO_PEER_MERGE") is None:
            try:
                import ctypes

                import torch.distributed._symmetric_memory as symm

                from .engine import CTN_MAILBOX_WORDS

                grp = group if group is not None else dist.group.WORLD
                box = symm.empty(CTN_MAILBOX_WORDS, dtype=torch.float64, device=dev)
                box.zero_()
                hdl = symm.rendezvous(box, grp)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                assert len(ptrs) == world and all(ptrs)
                torch.cuda.synchronize(dev)
                dist.barrier(group=group)  # every mailbox is zeroed before anybody posts
                st = {"box": box, "hdl": hdl, "ptrs": (ctypes.c_void_p * world)(*ptrs), "seq": 0,
                      "c_out": torch.empty(1, dtype=torch.float64, device=dev), "group": grp, "world": world,
                      "arena": None}
            except Exception as exc:  # no symmetric memory on this platform / build
                import warnings

                warnings.warn(f"peer mailbox merge unavailable ({type(exc).__name__}: {exc}); using an NCCL all-gather")
                st = False
        _PEER[key] = st
    return st or None


def peer_arena(st, nbytes):
    """Peer-mapped receive arena of at least ``nbytes`` for the one-kernel re-sharding
    (``ctn_reshard_peer``), plus the second mailbox through which the ranks report that their stores
    have landed and the kernel's ticket word.  COLLECTIVE: every rank of the group calls it with the
    same size.  Returns ``(arena tensor (uint8), [base pointer per rank])``."""
    import ctypes

    import torch.distributed._symmetric_memory as symm

    from .engine import CTN_MAILBOX_WORDS

    dev, world, grp = st["box"].device, st["world"], st["group"]
    if "done" not in st:
        done = symm.empty(CTN_MAILBOX_WORDS, dtype=torch.float64, device=dev)
        done.zero_()
        hdl = symm.rendezvous(done, grp)
        st["done"], st["done_hdl"] = done, hdl
        st["done_ptrs"] = (ctypes.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
        st["ticket"] = torch.zeros(4, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        dist.barrier(group=grp)
    if st["arena"] is None or st["arena"][0].numel() < nbytes:
        arena = symm.empty(int(nbytes), dtype=torch.uint8, device=dev)
        hdl = symm.rendezvous(arena, grp)
        st["arena"] = (arena, [int(p) for p in hdl.buffer_ptrs], hdl)
        torch.cuda.synchronize(dev)
        dist.barrier(group=grp)
    return st["arena"][0], st["arena"][1]


def reshard_peer(send, plan, st, arena_offset, group=None):
    """ONE kernel: common-scale merge of the ranks' slabs and all-to-all over peer memory.  ``send`` is
    this rank's slab ``[world][chunk]`` as ``plan`` (possibly a ``CTN_PLAN_RAW_FINAL`` one) left it; the
    result is this rank's ``[world][chunk]`` receive buffer at ``arena_offset`` of the peer arena and the
    common log-scale (a 1-element tensor that the next call overwrites)."""
    import ctypes

    from . import engine

    lib = engine.load_library()
    world = st["world"]
    rank = dist.get_rank(group)
    arena, bases = st["arena"][0], st["arena"][1]
    nbytes = send.numel() * send.element_size()
    assert arena_offset % 256 == 0 and arena_offset + nbytes <= arena.numel()
    recv_ptrs = (ctypes.c_void_p * world)(*[b + arena_offset for b in bases])
    dt = engine.CTN_F32 if send.dtype == torch.float32 else engine.CTN_F64
    st["seq"] += 1
    stream = torch.cuda.current_stream(send.device).cuda_stream
    engine._check(lib.ctn_reshard_peer(send.data_ptr(), send.numel(), dt, plan._dev[2], plan.final_kappa_offset,
                                       st["ptrs"], st["done_ptrs"], recv_ptrs, st["ticket"].data_ptr(), world, rank,
                                       st["seq"], st["c_out"].data_ptr(), stream), "ctn_reshard_peer")
    recv = arena[arena_offset:arena_offset + nbytes].view(send.dtype).view(send.shape)
    return recv, st["c_out"]


def merge_batch_shards(T_local, c_local, group=None, plan=None):
    """Global split-format pair for a batch-sharded evaluation.

    ``T_local`` is this rank's shard of the output, ``c_local`` its log-scale: a
    0-d/1-element float64 tensor (only element 0 is read).  Returns
    ``(T_shard, c)`` with ONE log-scale ``c`` shared by all ranks such that the
    concatenation of the shards has mean-abs 1, exactly what the reference
    returns for the unsharded network.

    CUDA, with the :class:`~contractn_b200.engine.Plan` that just produced ``T_local``: ONE kernel
    (``ctn_merge_batch_shards_peer``) exchanges the ranks' (log-scale, sum|T|, numel) triples through
    peer-mapped mailboxes over NVLink and rescales the shard -- no host-launched collective, the triple
    comes from the plan's own state block.  Otherwise: one small all-gather (3 doubles per rank).
    """
    if _world(group) == 1:
        return T_local, c_local.reshape(-1)[:1].to(torch.float64)
    world = _world(group)
    if T_local.is_cuda and T_local.is_contiguous() and T_local.dtype in (torch.float32, torch.float64) \
            and c_local.dtype == torch.float64:
        from . import engine

        lib = engine.load_library()
        dev = T_local.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        dt = engine.CTN_F32 if T_local.dtype == torch.float32 else engine.CTN_F64
        rank = dist.get_rank(group)
        ws_ptr = plan._dev[2] if plan is not None and getattr(plan, "_dev", None) is not None else None
        peer = _peer_state(dev, world, group) if ws_ptr is not None else None
        if peer is not None:
            peer["seq"] += 1
            engine._check(lib.ctn_merge_batch_shards_peer(T_local.data_ptr(), T_local.numel(), dt, ws_ptr, peer["ptrs"],
                                                          world, rank, peer["seq"], peer["c_out"].data_ptr(), stream),
                          "ctn_merge_batch_shards_peer")
            return T_local, peer["c_out"]
        # NCCL path: all-gather of 24 B per rank, then the rescale kernel
        key = (dev.index, world)
        if key not in _TABLES:
            _TABLES[key] = (torch.empty(world * 3, dtype=torch.float64, device=dev),
                            torch.empty(3, dtype=torch.float64, device=dev))
        table, mine = _TABLES[key]
        mine[0] = c_local.reshape(-1)[0]
        mine[1] = torch.linalg.vector_norm(T_local.reshape(-1), ord=1).to(torch.float64)
        mine[2] = float(T_local.numel())
        dist.all_gather_into_tensor(table, mine, group=group)
        c_out = torch.empty(1, dtype=torch.float64, device=dev)
        engine._check(lib.ctn_merge_batch_shards(T_local.data_ptr(), T_local.numel(), dt, table.data_ptr(), world,
                                                 rank, c_out.data_ptr(), stream),
                      "ctn_merge_batch_shards")
        return T_local, c_out
    c_local = c_local.reshape(-1)[0].to(torch.float64)
    # sum|T_local| is numel when the shard was rescaled by its last step, but a
    # shard whose norms fall under the reference's 1e-7 threshold is returned
    # unscaled (contractn/einsum.py:102), so measure it (one reduction pass)
    abs_sum = torch.linalg.vector_norm(T_local.reshape(-1), ord=1).to(torch.float64)
    mine = torch.stack([c_local, abs_sum, torch.full_like(c_local, float(T_local.numel()))])
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)  # (gloo has no all_gather_into_tensor)
    table = torch.stack(parts)
    c_star = table[:, 0].max()
    w_all = torch.exp(table[:, 0] - c_star)
    kappa = (table[:, 1] * w_all).sum() / table[:, 2].sum()
    w = torch.exp(c_local - c_star)
    T = T_local * (w / kappa).to(T_local.dtype)
    return T, (c_star + torch.log(kappa)).reshape(1)


def lse_accumulate(acc, T, c):
    """Running sum of split-format terms on one rank: ``acc`` is ``None`` or a
    pair ``(M, s)`` meaning ``M * exp(s)``; returns the pair including ``T*exp(c)``."""
    c = c.reshape(-1)[0].to(torch.float64)
    if acc is None:
        return T.clone(), c.clone()
    M, s = acc
    top = torch.maximum(s, c)
    M = M * torch.exp(s - top).to(M.dtype) + T * torch.exp(c - top).to(T.dtype)
    return M, top


def merge_slice_sums(M_local, s_local, group=None):
    """Sum over ranks of ``M_r * exp(s_r)`` returned in canonical split format
    ``(T, c)`` with mean |T| == 1 (all ranks get the full result)."""
    s_local = s_local.reshape(-1)[0].to(torch.float64)
    if _world(group) > 1:
        s_star = s_local.clone()
        dist.all_reduce(s_star, op=dist.ReduceOp.MAX, group=group)
        M = M_local * torch.exp(s_local - s_star).to(M_local.dtype)
        dist.all_reduce(M, op=dist.ReduceOp.SUM, group=group)
    else:
        s_star, M = s_local, M_local
    norm = M.abs().sum().to(torch.float64)
    numel = max(1, M.numel())
    if float(norm) > 1e-7:  # same threshold as stabilize (contractn/einsum.py:94-102)
        kappa = norm / numel
        return M / kappa.to(M.dtype), (s_star + torch.log(kappa)).reshape(1)
    return M, s_star.reshape(1)


def slice_assignments(n_slices, rank=None, world=None):
    """Round-robin slice indices of this rank."""
    if world is None:
        world = _world()
    if rank is None:
        rank = dist.get_rank() if world > 1 else 0
    return list(range(rank, n_slices, world))


# ---------------------------------------------------------------------------
# bond slicing (host-side bookkeeping)
# ---------------------------------------------------------------------------
def _parse(einstr):
    ins, out = einstr.split("->")
    return ins.split(","), out


def _sliced_flops(step_labels, size, chosen):
    """Path FLOPs of ONE slice (opt_einsum convention) with `chosen` labels fixed."""
    total = 0
    for all_labels, summed in step_labels:
        vol = 1
        for s in all_labels:
            if s not in chosen:
                vol *= size[s]
        total += vol * (2 if any(s not in chosen for s in summed) else 1)
    return total


def choose_slice_labels(einstr, shapes, contraction_list, min_slices):
    """Pick contracted bonds to slice until there are at least ``min_slices``
    slices.  Greedy and path-aware: at every round take the label whose removal
    leaves the cheapest sliced path (total work = slices x per-slice FLOPs of the
    SAME pairwise sequence), i.e. the bond that lives in the most expensive
    steps -- "index slicing of the largest contracted bonds"."""
    terms, out = _parse(einstr)
    size = {}
    for t, shp in zip(terms, shapes):
        for s, d in zip(t, shp):
            size[s] = max(size.get(s, 1), int(d))
    step_labels = []
    for c in contraction_list:
        ins = c[2].split("->")[0]
        step_labels.append((set(ins.replace(",", "")), set(c[1])))
    candidates = [s for s in size if s not in out and size[s] > 1]
    chosen, n = [], 1
    while n < min_slices and candidates:
        best, best_cost = None, None
        for s in candidates:
            cost = _sliced_flops(step_labels, size, set(chosen) | {s}) * n * size[s]
            if best_cost is None or cost < best_cost or (cost == best_cost and s < best):
                best, best_cost = s, cost
        chosen.append(best)
        candidates.remove(best)
        n *= size[best]
    return tuple(chosen)


class SlicedNetwork:
    """An einsum network with some contracted labels fixed per slice.

    ``einstr`` is the network with the sliced symbols deleted from every term;
    ``slice_operands(ops, k)`` returns views of the operands for slice ``k``
    (no copies: basic indexing, the engine consumes the strides)."""

    def __init__(self, einstr, shapes, labels):
        terms, out = _parse(einstr)
        for s in labels:
            if s in out:
                raise ValueError(f"cannot slice output label {s!r}")
        self.labels = tuple(labels)
        self.extents = []
        for s in labels:
            ext = [shp[t.index(s)] for t, shp in zip(terms, shapes) if s in t]
            if not ext:
                raise ValueError(f"label {s!r} not in the network")
            self.extents.append(int(max(ext)))
        self.n_slices = int(math.prod(self.extents)) if labels else 1
        self.terms = terms
        self.einstr = ",".join("".join(c for c in t if c not in labels) for t in terms) + "->" + out
        self.shapes = tuple(tuple(d for c, d in zip(t, shp) if c not in labels) for t, shp in zip(terms, shapes))

    def slice_index(self, k):
        idx = []
        for e in reversed(self.extents):
            idx.append(k % e)
            k //= e
        return tuple(reversed(idx))

    def slice_operands(self, operands, k):
        where = dict(zip(self.labels, self.slice_index(k)))
        out = []
        for t, op in zip(self.terms, operands):
            sel = tuple(where[c] if c in where else slice(None) for c in t)
            out.append(op[sel] if any(c in where for c in t) else op)
        return out


def contract_sliced(einstr, *operands, slice_labels=None, min_slices=None, optimize="auto",
                    split_format=True, group=None, contract_fn=None):
    """``contract`` with the largest bonds sliced across the ranks of ``group``.

    Every rank passes the SAME full operands (they are small next to the
    intermediates: this is the regime where slicing is used); rank r contracts
    slices r, r+W, ... on its own GPU and the partial sums are merged with an
    all-reduce of (mantissa, log-scale) pairs.  Returns the reference's result
    for the unsliced network, on every rank.
    """
    from . import planner
    from .engine import contract as engine_contract

    contract_fn = contract_fn or engine_contract
    shapes = tuple(tuple(op.shape) for op in operands)
    world = _world(group)
    rank = dist.get_rank(group) if world > 1 else 0
    if slice_labels is None:
        lst = planner.cached_contract_list(einstr, shapes, optimize=optimize)
        slice_labels = choose_slice_labels(einstr, shapes, lst, min_slices or world)
    net = SlicedNetwork(einstr, shapes, slice_labels)
    acc = None
    for k in slice_assignments(net.n_slices, rank, world):
        T, c = contract_fn(net.einstr, *net.slice_operands(operands, k), optimize=optimize, split_format=True)
        T = torch.as_tensor(T)
        c = torch.as_tensor(c, dtype=torch.float64, device=T.device)
        acc = lse_accumulate(acc, T, c)
    if acc is None:  # more ranks than slices: contribute an exact zero
        T0, _ = contract_fn(net.einstr, *net.slice_operands(operands, 0), optimize=optimize, split_format=True)
        T0 = torch.as_tensor(T0)
        acc = (torch.zeros_like(T0), torch.full((), -1.0e300, dtype=torch.float64, device=T0.device))
    T, c = merge_slice_sums(acc[0], acc[1], group=group)
    if split_format:
        return T, c
    return T * torch.exp(c).to(T.dtype)


# ---------------------------------------------------------------------------
# output-slab sharding (cfg 4: the output is far bigger than the operands)
# ---------------------------------------------------------------------------
def slab_bounds(extent, rank, world):
    """Contiguous, near-equal split of an open index of size ``extent``."""
    return rank * extent // world, (rank + 1) * extent // world


def contract_output_slabs(einstr, *operands, label=None, optimize="auto", group=None, contract_fn=None,
                          merge_fn=None):
    """``contract(..., split_format=True)`` with ONE OPEN index split across the ranks of ``group``.

    An output entry depends only on the operand rows with the same value of an open index, so every rank
    contracts the network with that index restricted to its slab -- views of the operands, no copies, no
    exchange of tensor data -- and keeps its slab of the result: a CP reconstruction whose output (550 GB
    at BASELINE's literal size) fits no single GPU is held slab-wise (cfg 4).  The
    split-format pair needs ONE global scale: the same 24-byte exchange as for batch shards
    (:func:`merge_batch_shards`).  Returns ``(T_slab, c, (lo, hi))``; ``label`` defaults to the first
    output index.  Contracted-rank slicing would need an all-reduce of the whole output instead.
    """
    from .engine import contract as engine_contract

    contract_fn = contract_fn or engine_contract
    terms, out = _parse(einstr)
    if not out:
        raise ValueError("the network has no open index to shard")
    label = label or out[0]
    if label not in out:
        raise ValueError(f"{label!r} is not an output index of {einstr!r}")
    world = _world(group)
    rank = dist.get_rank(group) if world > 1 else 0
    extent = max(op.shape[t.index(label)] for t, op in zip(terms, operands) if label in t)
    lo, hi = slab_bounds(int(extent), rank, world)
    sliced = []
    for t, op in zip(terms, operands):
        if label in t:
            sel = tuple(slice(lo, hi) if c == label else slice(None) for c in t)
            op = op[sel]
        sliced.append(op)
    res = contract_fn(einstr, *sliced, optimize=optimize, split_format=True)
    T = torch.as_tensor(res[0])
    c = torch.as_tensor(res[1], dtype=torch.float64, device=T.device).reshape(1)
    if merge_fn is not None:
        T, c = merge_fn(T, c)
    else:
        T, c = merge_batch_shards(T, c, group=group)
    return T, c, (lo, hi)
