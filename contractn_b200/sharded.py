"""Boundary-sharded sweep: bond-parallel contraction that does NOT multiply the work.

``parallel.contract_sliced`` fixes a bond to one value per slice and re-runs the same
pairwise sequence per slice; on a 2-D grid a bond lives in the boundary for only L of
the L^2 sweep steps, so 8 slices cost 6.65x the FLOPs of the unsliced contraction
(DESIGN.md section 7).  Here the running intermediate of the path (the boundary tensor
of a sweep) is SHARDED over one of its legs instead:

* a step that does not sum the sharded leg ``s`` is purely local -- every rank contracts
  its slab of the boundary with (views of) the replicated small operands: 1/W of the
  FLOPs per rank, no communication;
* just before the step that sums ``s`` the boundary is re-sharded onto the leg ``s'``
  that stays open the longest: ONE all-to-all.  The producing segment writes its output
  in the layout ``(s'_hi, s_lo, s'_lo, rest)`` (``s'`` is split into two einsum labels
  by *viewing* the operands that carry it), so the send buffer is already ordered by
  destination rank and the receive buffer ``(s_hi, s_lo, s'_lo, rest)`` IS the new
  local tensor ``(s, s'_lo, rest)``: no pack or unpack pass on either side;
* when the boundary has shrunk under ``min_numel`` it is all-gathered and the tail of
  the path runs replicated.

Each run of consecutive local steps is ONE device plan (``engine.Plan``), so the
stabilisation of the reference (``contractn/einsum.py:89-107``, rescale after every
pairwise step, log-scales summed) happens inside the kernels as everywhere else; a
rank's slab carries its OWN scale, and the slabs are put on one common scale by the
one-kernel peer-mailbox merge (``parallel.merge_batch_shards``) before they are mixed
by an exchange.  ``T * exp(c)`` is what the reference computes; the result pair is
canonical (mean |T| == 1) because the last step is a stabilised one.

The host logic (segmentation, label bookkeeping, layouts) is independent of the device:
``contract_fn`` lets the CPU tests run it over gloo with the oracle doing the segments.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from . import planner
from .parallel import _parse, _world, merge_batch_shards

__all__ = ["ShardedSweep", "contract_sharded_sweep", "plan_segments"]

_INF = 1 << 60


class _Segment:
    """A run of consecutive path steps along the running intermediate, executed as one plan."""

    def __init__(self, mode, s=None):
        self.mode = mode            # "repl": every rank holds everything; "shard": the chain is sharded over `s`
        self.s = s
        self.steps = []             # [(input ids, output id)]
        self.chain = None           # id of the last step's output
        self.chain_labels = ""      # its labels (original order of the path's einsum strings)
        self.exchange = None        # None | "a2a" | "gather" | "final"
        self.s2 = None              # next shard label (a2a)
        self.hi = None              # fresh symbol for the high part of s2 (a2a)
        self.out_id = None          # id under which the result (after the exchange) is stored
        self.chain_in = None        # id of the sharded/running input (listed LAST, like the reference's sweep paths)
        self.inputs = []            # external input ids, in order of first use (the chain input last)
        self.terms = []             # their labels as this segment's einsum sees them
        self.views = []             # per input: (slice_dim or None, split_dim or None)
        self.out = ""               # output labels of the segment's einsum
        self.post_labels = ""       # labels of the stored result as later segments see it
        self.post_local = None      # label the stored result is sharded over (None: replicated)
        self.path = ()              # linear path of the segment's sub-network

    def einstr(self):
        return ",".join(self.terms) + "->" + self.out


def _fresh_symbols(used):
    i = 0
    while True:
        ch = planner.get_symbol(i)
        i += 1
        if ch not in used:
            used.add(ch)
            yield ch


def plan_segments(einstr, shapes, contraction_list, world, min_numel=1 << 16):
    """Host-side planning, identical on every rank.  Returns ``(segments, n_inputs, out_perm)``:
    ``out_perm`` permutes the stored final result into the network's output order (``None`` when it
    already is)."""
    terms, out = _parse(einstr)
    size = {}
    for t, shp in zip(terms, shapes):
        for ch, d in zip(t, shp):
            size[ch] = max(size.get(ch, 1), int(d))
    n = len(terms)
    live, labels, nxt = list(range(n)), dict(enumerate(terms)), n
    steps = []
    for c in contraction_list:
        pos = list(c[0])
        ins, outs = c[2].split("->")
        ids = [live[p] for p in pos]
        assert [set(labels[i]) for i in ids] == [set(t) for t in ins.split(",")], "path/einsum mismatch"
        for p in sorted(pos, reverse=True):
            live.pop(p)
        live.append(nxt)
        labels[nxt] = outs
        steps.append((tuple(ids), nxt, set(c[1])))
        nxt += 1
    gone = {}                       # label -> index of the step that sums it (absent: open index)
    for t, (_, _, removed) in enumerate(steps):
        for ch in removed:
            gone.setdefault(ch, t)
    fresh = _fresh_symbols(set(einstr))
    local = {}                      # id -> label the stored tensor is sharded over

    def numel(lab):
        return math.prod(size[ch] for ch in lab)

    def pick(lab, exclude):
        cands = [ch for ch in lab if ch not in exclude and size[ch] >= world and size[ch] % world == 0]
        if not cands:
            return None
        return max(cands, key=lambda ch: (gone.get(ch, _INF), -lab.index(ch)))

    segments, pending = [], None

    def close(seg, exchange=None, s2=None):
        nonlocal nxt
        seg.exchange, seg.s2 = exchange, s2
        made = {o for _, o in seg.steps}
        for ids, _ in seg.steps:
            for i in ids:
                if i not in made and i not in seg.inputs:
                    seg.inputs.append(i)
        if seg.chain_in in seg.inputs:
            # the running tensor goes last: the positional path then reads (k, 0), (k, 0), ... and the
            # planner's step strings put the big operand first, as in the unsharded sweep
            seg.inputs.remove(seg.chain_in)
            seg.inputs.append(seg.chain_in)
        if exchange == "a2a":
            seg.hi = next(fresh)
        for i in seg.inputs:
            lab, sl, sp = labels[i], None, None
            if seg.mode == "shard" and seg.s in lab and local.get(i) != seg.s:
                sl = lab.index(seg.s)
            if exchange == "a2a" and s2 in lab:
                sp = lab.index(s2)
                lab = lab.replace(s2, seg.hi + s2)
            seg.terms.append(lab)
            seg.views.append((sl, sp))
        rest = seg.chain_labels
        if exchange is None:
            seg.out, seg.post_labels, seg.out_id = rest, rest, seg.chain
        else:
            s = seg.s
            if exchange == "a2a":
                rest = "".join(ch for ch in rest if ch not in (s, s2))
                seg.out = seg.hi + s + s2 + rest
                seg.post_labels, seg.post_local = s + s2 + rest, s2
            else:
                rest = "".join(ch for ch in rest if ch != s)
                seg.out = seg.post_labels = s + rest
            seg.out_id = nxt
            nxt += 1
            labels[seg.out_id] = seg.post_labels
            if seg.post_local:
                local[seg.out_id] = seg.post_local
        ssa, sid, k = [], {i: j for j, i in enumerate(seg.inputs)}, len(seg.inputs)
        for ids, o in seg.steps:
            ssa.append(tuple(sid[i] for i in ids))
            sid[o] = k
            k += 1
        seg.path = tuple(planner._ssa_to_linear(ssa, len(seg.inputs)))
        segments.append(seg)
        return seg.out_id

    def start(ids, o, removed):
        lab = labels[o]
        s = pick(lab, removed) if numel(lab) >= min_numel else None
        # (a label that is summed by the very next chain step would be re-sharded at once: still correct)
        seg = _Segment("shard", s) if s is not None else _Segment("repl")
        seg.steps.append((ids, o))
        seg.chain, seg.chain_labels = o, lab
        return seg

    for ids, o, removed in steps:
        ids = tuple(ids)
        if pending is not None and pending.chain in ids:
            if pending.mode == "shard":
                if pending.s in removed or pending.s not in labels[o]:
                    lab = pending.chain_labels
                    s2 = pick(lab, removed | {pending.s}) if numel(lab) >= min_numel else None
                    old = pending.chain
                    new = close(pending, "a2a" if s2 is not None else "gather", s2)
                    ids = tuple(new if i == old else i for i in ids)
                    pending = _Segment("shard", s2) if s2 is not None else _Segment("repl")
                    pending.chain_in = new
                pending.steps.append((ids, o))
                pending.chain, pending.chain_labels = o, labels[o]
            else:
                lab = labels[o]
                s = pick(lab, removed) if numel(lab) >= min_numel else None
                if s is not None:
                    prev = pending.chain
                    close(pending)
                    pending = start(ids, o, removed)
                    pending.chain_in = prev
                else:
                    pending.steps.append((ids, o))
                    pending.chain, pending.chain_labels = o, lab
        elif pending is not None and pending.mode == "shard":
            side = _Segment("repl")     # a step off the chain: replicated, emitted ahead of the pending segment
            side.steps.append((ids, o))
            side.chain, side.chain_labels = o, labels[o]
            close(side)
        else:
            if pending is not None:
                close(pending)
            pending = start(ids, o, removed)
    out_perm = None
    if pending is not None:
        if pending.mode == "shard":
            close(pending, "final")
            got = pending.post_labels
            if got != out:
                out_perm = tuple(got.index(ch) for ch in out)
        else:
            close(pending)
    return segments, n, out_perm


class ShardedSweep:
    """``contract(einstr, *operands, optimize=path, split_format=True)`` with the running intermediate
    sharded across the ranks of ``group``.  Every rank passes the SAME full operands (the site tensors of
    a grid are tiny next to the boundary).  ``run()`` may be called repeatedly on the same resident
    operands (their contents may change): plans, output and exchange buffers are set up once.

    ``contract_fn(einstr, *ops, optimize=path, split_format=True) -> (T, c)`` replaces the device plans
    (CPU tests); with it, tensors live wherever the operands live.

    Limits: extents of the sharded legs must be multiples of the world size (legs that are not are never
    picked; a network without such a leg runs replicated); the result flags of the segment plans are not
    read back, so a float32 plan with fused steps that meets a degenerate norm (``CTN_RESULT_DEGENERATE``,
    see ``engine._run_torch_nograd``) is not re-run in exact-step mode here -- the path this class exists
    for (float64 grids) has no fused steps.
    """

    def __init__(self, einstr, operands, optimize="auto", group=None, min_numel=1 << 16, contract_fn=None):
        self.group = group
        self.world = _world(group)
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.operands = [torch.as_tensor(o) for o in operands]
        shapes = tuple(tuple(o.shape) for o in self.operands)
        lst = planner.cached_contract_list(einstr, shapes, optimize=optimize)
        self.segments, self.n_inputs, self.out_perm = plan_segments(einstr, shapes, lst, self.world, min_numel)
        self.contract_fn = contract_fn
        self._bound = {}
        self._bufs = {}
        self.dtype = torch.float64 if any(o.dtype == torch.float64 for o in self.operands) else torch.float32
        self.exchanged_bytes = 0
        # CUDA, W > 1: the all-to-all exchanges are ONE kernel each over peer memory (scale merge fused, the
        # producing plan's finalize pass skipped); otherwise merge kernel + NCCL/gloo all_to_all_single
        self._peer = None
        if self.world > 1 and contract_fn is None and self.operands[0].is_cuda:
            from . import parallel

            self._peer = parallel._peer_state(self.operands[0].device, self.world, group)
            if self._peer is not None:
                esz = 8 if self.dtype == torch.float64 else 4
                size = {}
                for t, shp in zip(_parse(einstr)[0], shapes):
                    size.update(zip(t, shp))
                need = [math.prod(self.world if ch == s.hi else size[ch] for ch in s.out) // self.world ** 2 * esz
                        for s in self.segments if s.exchange == "a2a"]
                # two halves, used alternately (one would do: a peer stores into my buffer only after my
                # triple for THAT exchange is out, i.e. after I consumed the previous one)
                self._half = (max(need, default=0) + 255) // 256 * 256
                if need:
                    parallel.peer_arena(self._peer, 2 * self._half)

    # -- one segment ---------------------------------------------------------------------------
    def _views(self, seg, store):
        ops = []
        for i, (sl, sp) in zip(seg.inputs, seg.views):
            t = store[i]
            if sl is not None:
                e = t.shape[sl] // self.world
                t = t.narrow(sl, self.rank * e, e)
            if sp is not None:
                t = t.unflatten(sp, (self.world, t.shape[sp] // self.world))
            ops.append(t)
        return ops

    def _run_segment(self, k, seg, ops):
        """-> (out, log-scale as a 1-element float64 tensor, plan or None)"""
        if self.contract_fn is not None:
            T, c = self.contract_fn(seg.einstr(), *ops, optimize=seg.path if seg.path else "auto",
                                    split_format=True)
            T = torch.as_tensor(T)
            return T.contiguous(), torch.as_tensor(c, dtype=torch.float64).reshape(-1)[:1].clone(), None
        from . import engine
        import numpy as np

        key = (k, tuple(t.data_ptr() for t in ops))
        hit = self._bound.get(k)
        if hit is None or hit[0] != key:
            ops = [t if t.dtype == self.dtype else t.to(self.dtype) for t in ops]
            shapes = tuple(tuple(t.shape) for t in ops)
            strides = tuple(tuple(t.stride()) for t in ops)
            lst = planner.cached_contract_list(seg.einstr(), shapes, optimize=seg.path if seg.path else "auto")
            dev = ops[0].device
            raw = engine.CTN_PLAN_RAW_FINAL if seg.exchange == "a2a" and self._peer is not None else 0
            plan = engine.get_plan(lst, shapes, np.float64 if self.dtype == torch.float64 else np.float32,
                                   strides=strides, flags=raw, device=dev.index)
            out = torch.empty(plan.out_shape, dtype=self.dtype, device=dev)
            hit = (key, plan.bind(ops, out=out), plan)
            self._bound[k] = hit
        out, tail = hit[1].run()
        return out, tail[0:1], hit[2]

    def _buffer(self, k, shape, like):
        buf = self._bufs.get(k)
        if buf is None or buf.shape != tuple(shape):
            buf = torch.empty(shape, dtype=like.dtype, device=like.device)
            self._bufs[k] = buf
        return buf

    # -- the whole path --------------------------------------------------------------------------
    def run(self):
        store = dict(enumerate(self.operands))
        c_acc = None
        W = self.world
        self.exchanged_bytes = 0
        out_id, n_a2a = None, 0
        for k, seg in enumerate(self.segments):
            out, c, plan = self._run_segment(k, seg, self._views(seg, store))
            fused = seg.exchange == "a2a" and self._peer is not None
            if fused:
                from . import parallel

                n_a2a += 1
                recv, c = parallel.reshard_peer(out, plan, self._peer, (n_a2a & 1) * self._half, group=self.group)
                self.exchanged_bytes += out.numel() * out.element_size() * (W - 1) // W
            elif seg.exchange is not None and W > 1:
                out, c = merge_batch_shards(out, c, group=self.group, plan=plan)
            c_acc = c.clone() if c_acc is None else c_acc + c
            if seg.exchange == "a2a":
                if not fused:
                    recv = self._buffer(k, out.shape, out)
                    if W > 1:
                        dist.all_to_all_single(recv, out, group=self.group)
                        self.exchanged_bytes += out.numel() * out.element_size() * (W - 1) // W
                    else:
                        recv.copy_(out)
                # (s_hi, s_lo, s2_lo, rest) -> (s, s2_lo, rest): adjacent dims of a contiguous buffer
                out = recv.reshape((recv.shape[0] * recv.shape[1],) + tuple(recv.shape[2:]))
            elif seg.exchange in ("gather", "final") and W > 1:
                recv = self._buffer(k, (W * out.shape[0],) + tuple(out.shape[1:]), out)
                if out.is_cuda:
                    dist.all_gather_into_tensor(recv, out, group=self.group)
                else:  # gloo
                    parts = list(recv.chunk(W, dim=0))
                    dist.all_gather(parts, out, group=self.group)
                self.exchanged_bytes += out.numel() * out.element_size() * (W - 1)
                out = recv
            store[seg.out_id] = out
            out_id = seg.out_id
            for i in seg.inputs:    # intermediates die with their consumer
                if i >= self.n_inputs:
                    store.pop(i, None)
        T = store[out_id]
        if self.out_perm is not None:
            T = T.permute(self.out_perm)
        return T, c_acc

    def describe(self):
        lines = []
        for k, seg in enumerate(self.segments):
            lines.append(f"seg {k}: {seg.mode}{'(' + seg.s + ')' if seg.s else ''} {len(seg.steps)} steps "
                         f"{seg.einstr()} -> {seg.exchange or '-'}{'(' + seg.s2 + ')' if seg.s2 else ''}")
        return "\n".join(lines)


def contract_sharded_sweep(einstr, *operands, optimize="auto", split_format=True, group=None,
                           min_numel=1 << 16, contract_fn=None):
    """One-shot form of :class:`ShardedSweep` (mirrors ``contract``'s keywords)."""
    T, c = ShardedSweep(einstr, operands, optimize=optimize, group=group, min_numel=min_numel,
                        contract_fn=contract_fn).run()
    if split_format:
        return T, c
    return T * torch.exp(c.reshape(-1)[0]).to(T.dtype)
