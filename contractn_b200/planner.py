"""Host-side contraction-path planner (the one piece that stays on the CPU).

The reference obtains its pairwise contraction list from the third-party
package ``opt_einsum`` (``contractn/einsum.py:313-323`` calls
``oe.contract_path(einstr, *shapes, shapes=True, einsum_call=True, ...)``).
``opt_einsum`` is neither vendored in the reference nor installable in this
image (no wheel, no index), so this module

* uses the real ``opt_einsum`` when it can be imported at run time, and
* otherwise provides a from-scratch planner with the same call contract and
  the same published algorithms (``opt_einsum`` 3.3.0, the version floor in the
  reference's ``requirements.txt:2``): ``optimal``, ``branch-all`` /
  ``branch-2`` / ``branch-1``, ``greedy``, ``auto`` and explicit paths, plus
  the ``can_blas`` classifier that yields the 5th field of every contraction
  tuple (read at ``contractn/einsum.py:342,347``).

Path identity with the real library is NOT pinned by any reference test (see
SURVEY.md section 8c); values do not depend on it (finding 0.4).  The same
list is always handed to both the CPU oracle and the GPU engine.

Contraction list format (what ``_core_contract`` consumes,
``contractn/einsum.py:341-342``)::

    (positions_descending, idx_removed:set, "lhs,rhs->out", remaining|None, blas_flag)
"""
from __future__ import annotations

import heapq
import itertools
from collections import defaultdict
from functools import lru_cache

__all__ = [
    "get_symbol",
    "contract_path",
    "can_blas",
    "planner_name",
    "flop_count",
    "path_flops",
]

_LETTERS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"

try:  # pragma: no cover - not available in the build image
    import opt_einsum as _real_oe
except Exception:  # ModuleNotFoundError in this image
    _real_oe = None


def planner_name():
    """Which planner produced the paths: reported next to every benchmark."""
    if _real_oe is not None:
        return "opt_einsum-" + getattr(_real_oe, "__version__", "unknown")
    return "contractn_b200.planner (opt_einsum 3.3.0 restatement)"


def get_symbol(i):
    """Index -> einsum symbol; inverse is ``contractn/utils.py:67-81``."""
    if i < 52:
        return _LETTERS[i]
    return chr(i + 140)


# --------------------------------------------------------------------------
# size / cost helpers
# --------------------------------------------------------------------------
def _size_of(indices, sizes):
    n = 1
    for s in indices:
        n *= sizes[s]
    return n


def flop_count(all_indices, has_inner, n_terms, sizes):
    """opt_einsum / numpy FLOP convention: prod(extents) * (n_terms-1 [+1 if summed])."""
    factor = max(1, n_terms - 1)
    if has_inner:
        factor += 1
    return _size_of(all_indices, sizes) * factor


def _parse(subscripts, shapes):
    subscripts = subscripts.replace(" ", "")
    if "." in subscripts:
        raise ValueError("ellipsis subscripts are not supported by this planner")
    if "->" in subscripts:
        lhs, out = subscripts.split("->")
    else:
        lhs = subscripts
        flat = lhs.replace(",", "")
        out = "".join(sorted(s for s in set(flat) if flat.count(s) == 1))
    terms = lhs.split(",")
    if len(terms) != len(shapes):
        raise ValueError(
            "Number of einsum subscripts must be equal to the number of operands."
        )
    for s in out:
        if s not in lhs:
            raise ValueError("Output character '{}' did not appear in the input".format(s))
    return terms, out


def _size_dict(terms, shapes):
    sizes = {}
    for pos, (term, shp) in enumerate(zip(terms, shapes)):
        if len(term) != len(shp):
            raise ValueError(
                "Einstein sum subscript '{}' does not contain the correct number "
                "of indices for operand {}.".format(term, pos)
            )
        for sym, dim in zip(term, shp):
            dim = int(dim)
            if sym in sizes:
                # broadcasting: the largest extent wins, 1 is compatible with all
                if sizes[sym] == 1:
                    sizes[sym] = dim
                elif dim not in (1, sizes[sym]):
                    raise ValueError(
                        "Size of label '{}' for operand {} ({}) does not match "
                        "previous terms ({}).".format(sym, pos, sizes[sym], dim)
                    )
            else:
                sizes[sym] = dim
    return sizes


def _memory_arg(memory_limit, size_list):
    if memory_limit == "max_input":
        return max(size_list)
    if memory_limit is None:
        return None
    if memory_limit < 1:
        if memory_limit == -1:
            return None
        raise ValueError("Memory limit must be larger than 0, or -1")
    return int(memory_limit)


_NO_LIMIT = (None, -1, float("inf"))


# --------------------------------------------------------------------------
# SSA <-> linear path conversion
# --------------------------------------------------------------------------
def _ssa_to_linear(ssa_path, n_inputs):
    """Convert static-single-assignment ids to pop/append positions."""
    alive = list(range(n_inputs))  # live ssa ids; list position == linear index
    nxt = n_inputs
    out = []
    for step in ssa_path:
        out.append(tuple(alive.index(s) for s in step))
        for s in step:
            alive.remove(s)
        alive.append(nxt)
        nxt += 1
    return out


def _pair_result(inputs, output, remaining, i, j, sizes):
    a, b = inputs[i], inputs[j]
    both = a | b
    keep = frozenset.union(output, *(inputs[r] for r in remaining - {i, j}))
    result = both & keep
    flops = flop_count(both, (a & b) - keep, 2, sizes)
    return result, flops


def _oversize_flops(inputs, remaining, output, sizes):
    everything = frozenset.union(*(inputs[r] for r in remaining))
    return flop_count(everything, everything - output, len(remaining), sizes)


# --------------------------------------------------------------------------
# optimal: exhaustive depth-first search over pair orders
# --------------------------------------------------------------------------
def _optimal(inputs, output, sizes, memory_limit=None):
    inputs = tuple(frozenset(x) for x in inputs)
    output = frozenset(output)
    best = {"flops": float("inf"), "ssa": (tuple(range(len(inputs))),)}
    pair_cache = {}
    size_cache = {}

    def descend(path, live, tensors, flops):
        if len(live) == 1:
            best["flops"], best["ssa"] = flops, path
            return
        for i, j in itertools.combinations(live, 2):
            if i > j:
                i, j = j, i
            key = (tensors[i], tensors[j])
            if key not in pair_cache:
                pair_cache[key] = _pair_result(tensors, output, live, i, j, sizes)
            res, step_flops = pair_cache[key]
            total = flops + step_flops
            if total >= best["flops"]:
                continue
            if memory_limit not in _NO_LIMIT:
                if res not in size_cache:
                    size_cache[res] = _size_of(res, sizes)
                if size_cache[res] > memory_limit:
                    alt = flops + _oversize_flops(tensors, live, output, sizes)
                    if alt < best["flops"]:
                        best["flops"], best["ssa"] = alt, path + (tuple(live),)
                    continue
            descend(path + ((i, j),), live - {i, j} | {len(tensors)}, tensors + (res,), total)

    descend((), set(range(len(inputs))), inputs, 0)
    return _ssa_to_linear(best["ssa"], len(inputs))


# --------------------------------------------------------------------------
# branch-and-bound ("branch-all", "branch-2", "branch-1")
# --------------------------------------------------------------------------
def _branch(inputs, output, sizes, memory_limit=None, nbranch=None, cutoff_flops_factor=4):
    inputs = tuple(frozenset(x) for x in inputs)
    output = frozenset(output)
    best = {"flops": float("inf"), "size": float("inf"), "ssa": None}
    progress = defaultdict(lambda: float("inf"))
    size_cache = {k: _size_of(k, sizes) for k in inputs}
    pair_cache = {}

    def descend(path, tensors, live, flops, size):
        if len(live) == 1:
            best["size"], best["flops"], best["ssa"] = size, flops, path
            return

        def assess(a, b, i, j):
            if (a, b) not in pair_cache:
                pair_cache[a, b] = _pair_result(tensors, output, live, i, j, sizes)
            res, step_flops = pair_cache[a, b]
            if res not in size_cache:
                size_cache[res] = _size_of(res, sizes)
            res_size = size_cache[res]
            total = flops + step_flops
            peak = max(size, res_size)
            # minimise flops first, size second
            if not (total, peak) < (best["flops"], best["size"]):
                return None
            depth = len(tensors)
            if total < progress[depth]:
                progress[depth] = total
            elif total > cutoff_flops_factor * progress[depth]:
                return None
            if memory_limit not in _NO_LIMIT and res_size > memory_limit:
                alt = flops + _oversize_flops(tensors, live, output, sizes)
                if alt < best["flops"]:
                    best["flops"], best["ssa"] = alt, path + (tuple(live),)
                return None
            # "memory-removed" heuristic orders the candidates locally
            cost = res_size - size_cache[tensors[i]] - size_cache[tensors[j]]
            return cost, step_flops, total, peak, (i, j), res

        heap = []
        for outer_pass in (False, True):
            if outer_pass and heap:
                break
            for i, j in itertools.combinations(live, 2):
                if i > j:
                    i, j = j, i
                a, b = tensors[i], tensors[j]
                if not outer_pass and a.isdisjoint(b):
                    continue  # outer products only if nothing else is left
                cand = assess(a, b, i, j)
                if cand:
                    heapq.heappush(heap, cand)

        taken = 0
        while heap and (nbranch is None or taken < nbranch):
            _, _, total, peak, (i, j), res = heapq.heappop(heap)
            descend(path + ((i, j),), tensors + (res,), (live - {i, j}) | {len(tensors)}, total, peak)
            taken += 1

    descend((), inputs, set(range(len(inputs))), 0, 0)
    return _ssa_to_linear(best["ssa"], len(inputs))


# --------------------------------------------------------------------------
# greedy (ssa_greedy_optimize, cost = memory removed)
# --------------------------------------------------------------------------
def _greedy_ssa(inputs, output, sizes):
    if len(inputs) == 1:
        return [(0,)]
    inputs = [frozenset(x) for x in inputs]
    # an index shared by every operand cannot be summed before the last step:
    # treat it as an output (batch) index
    output = frozenset(output) | frozenset.intersection(*inputs)

    live = {}  # index set -> ssa id
    fresh = itertools.count(len(inputs))
    ssa = []
    # operands with identical index sets are folded eagerly (Hadamard products)
    for ident, key in enumerate(inputs):
        if key in live:
            ssa.append((live[key], ident))
            live[key] = next(fresh)
        else:
            live[key] = ident

    users = defaultdict(set)  # contractible index -> keys that carry it
    for key in live:
        for s in key - output:
            users[s].add(key)
    refs = {
        n: set(s for s, ks in users.items() if len(ks) >= n) - output for n in (2, 3)
    }
    foot = {key: _size_of(key, sizes) for key in live}

    def candidate(k1, k2):
        either, two = k1 | k2, k1 & k2
        one = either - two
        k12 = (either & output) | (two & refs[3]) | (one & refs[2])
        cost = _size_of(k12, sizes) - foot[k1] - foot[k2]
        id1, id2 = live[k1], live[k2]
        if id1 > id2:
            k1, id1, k2, id2 = k2, id2, k1, id1
        return (cost, id2, id1), k1, k2, k12

    heap = []
    for s, keys in users.items():
        keys = sorted(keys, key=live.__getitem__)
        for n, k1 in enumerate(keys[:-1]):
            heapq.heappush(heap, min(candidate(k1, k2) for k2 in keys[n + 1:]))

    while heap:
        _, k1, k2, k12 = heapq.heappop(heap)
        if k1 not in live or k2 not in live:
            continue  # stale candidate
        id1, id2 = live.pop(k1), live.pop(k2)
        for s in k1 - output:
            users[s].remove(k1)
        for s in k2 - output:
            users[s].remove(k2)
        ssa.append((id1, id2))
        if k12 in live:
            ssa.append((live[k12], next(fresh)))
        else:
            for s in k12 - output:
                users[s].add(k12)
        live[k12] = next(fresh)
        # (precedence as published: k1 | (k2 - output))
        for s in k1 | k2 - output:
            n = len(users[s])
            if n <= 1:
                refs[2].discard(s)
                refs[3].discard(s)
            elif n == 2:
                refs[2].add(s)
                refs[3].discard(s)
            else:
                refs[2].add(s)
                refs[3].add(s)
        foot[k12] = _size_of(k12, sizes)
        partners = set(k for s in k12 - output for k in users[s])
        partners.discard(k12)
        if partners:
            heapq.heappush(heap, min(candidate(k12, k2) for k2 in partners))

    # disconnected remainder: outer products, smallest first
    heap = [(_size_of(key & output, sizes), ident, key) for key, ident in live.items()]
    heapq.heapify(heap)
    _, id1, k1 = heapq.heappop(heap)
    while heap:
        _, id2, k2 = heapq.heappop(heap)
        ssa.append((min(id1, id2), max(id1, id2)))
        k12 = (k1 | k2) & output
        _, id1, k1 = heapq.heappushpop(heap, (_size_of(k12, sizes), next(fresh), k12))
    return ssa


def _greedy(inputs, output, sizes, memory_limit=None):
    if memory_limit not in _NO_LIMIT:
        return _branch(inputs, output, sizes, memory_limit, nbranch=1)
    return _ssa_to_linear(_greedy_ssa(inputs, output, sizes), len(inputs))


def _auto(inputs, output, sizes, memory_limit=None):
    n = len(inputs)
    if n <= 4:
        return _optimal(inputs, output, sizes, memory_limit)
    if n <= 6:
        return _branch(inputs, output, sizes, memory_limit, nbranch=None)
    if n <= 8:
        return _branch(inputs, output, sizes, memory_limit, nbranch=2)
    if n <= 14:
        return _branch(inputs, output, sizes, memory_limit, nbranch=1)
    return _greedy(inputs, output, sizes, memory_limit)


_OPTIMIZERS = {
    "auto": _auto,
    "optimal": _optimal,
    "greedy": _greedy,
    "eager": _greedy,
    "opportunistic": _greedy,
    "branch-all": lambda i, o, s, m=None: _branch(i, o, s, m, nbranch=None),
    "branch-2": lambda i, o, s, m=None: _branch(i, o, s, m, nbranch=2),
    "branch-1": lambda i, o, s, m=None: _branch(i, o, s, m, nbranch=1),
}


# --------------------------------------------------------------------------
# BLAS-ability classifier
# --------------------------------------------------------------------------
def can_blas(inputs, result, idx_removed, shapes=None):
    """Classify a pairwise step; the flag routes it at ``contractn/einsum.py:347``.

    Returns False (-> einsum branch) for anything that is not a plain two-operand
    tensor contraction: repeated index inside one operand, an index carried by
    both operands and the result (hyperedge / batch), an index summed out of one
    operand only, or a contracted index whose extents differ (broadcast).
    """
    if len(inputs) != 2:
        return False
    left, right = inputs
    for s in set(left + right):
        nl, nr = left.count(s), right.count(s)
        if nl > 1 or nr > 1 or nl + nr > 2:
            return False
        if nl + nr - 1 == int(s in result):
            return False
    if shapes is not None:
        for s in idx_removed:
            if shapes[0][left.find(s)] != shapes[1][right.find(s)]:
                return False
    if len(idx_removed) == 0:
        return "OUTER/EINSUM"
    set_l, set_r = set(left), set(right)
    keep_l, keep_r = set_l - idx_removed, set_r - idx_removed
    n = len(idx_removed)
    if left == right:
        return "DOT"
    if set_l == set_r:
        return "DOT/EINSUM"
    if left[-n:] == right[:n] or left[:n] == right[-n:]:
        return "GEMM"
    if left[-n:] == right[-n:] or left[:n] == right[:n]:
        return "GEMM"
    if not keep_l or not keep_r:
        return "GEMV/EINSUM"
    return "TDOT"


def _result_shape(terms, shapes, out):
    dims = {}
    for term, shp in zip(terms, shapes):
        for s, d in zip(term, shp):
            dims[s] = max(dims.get(s, 1), d)
    return tuple(dims[s] for s in out)


# --------------------------------------------------------------------------
# public entry point
# --------------------------------------------------------------------------
def _own_contract_path(subscripts, *shapes, optimize="auto", memory_limit=None, use_blas=True):
    terms, out = _parse(subscripts, shapes)
    shapes = [tuple(int(d) for d in s) for s in shapes]
    sizes = _size_dict(terms, shapes)
    index_sets = [set(t) for t in terms]
    out_set = set(out)
    size_list = [_size_of(t, sizes) for t in terms + [out]]
    mem = _memory_arg(memory_limit, size_list)
    n_ops = len(terms)

    if optimize is True:
        optimize = "auto"
    if not isinstance(optimize, str):
        path = [tuple(p) for p in optimize]  # explicit path
    elif n_ops <= 2:
        path = [tuple(range(n_ops))]
    else:
        try:
            finder = _OPTIMIZERS[optimize]
        except KeyError:
            raise KeyError(
                "Path optimizer '{}' is not implemented by the built-in planner "
                "(available: {}); install opt_einsum for the rest.".format(
                    optimize, sorted(_OPTIMIZERS)
                )
            )
        path = finder(index_sets, out_set, sizes, mem)

    terms = list(terms)
    contraction_list = []
    for n, positions in enumerate(path):
        positions = tuple(sorted(positions, reverse=True))
        # which indices survive this step
        popped = [index_sets.pop(p) for p in positions]
        touched = set().union(*popped)
        needed = out_set.union(*index_sets)
        new_set = needed & touched
        removed = touched - new_set
        index_sets.append(new_set)

        step_terms = [terms.pop(p) for p in positions]
        step_shapes = [shapes.pop(p) for p in positions]
        flag = can_blas(step_terms, new_set, removed, step_shapes) if use_blas else False
        if n == len(path) - 1:
            result = out
        else:
            joined = "".join(step_terms)
            result = "".join(sorted(new_set, key=joined.find))
        shapes.append(_result_shape(step_terms, step_shapes, result))
        terms.append(result)
        remaining = tuple(terms) if len(terms) <= 20 else None
        contraction_list.append(
            (positions, removed, ",".join(step_terms) + "->" + result, remaining, flag)
        )
    return contraction_list


def contract_path(subscripts, *shapes, **kwargs):
    """``opt_einsum.contract_path(..., shapes=True, einsum_call=True)`` work-alike.

    Returns ``(shapes, contraction_list)`` exactly as the reference unpacks it at
    ``contractn/einsum.py:322``.
    """
    allowed = {"optimize", "memory_limit", "einsum_call", "use_blas", "shapes"}
    unknown = [k for k in kwargs if k not in allowed]
    if unknown:
        raise TypeError("einsum_path: Did not understand the following kwargs: {}".format(unknown))
    if not kwargs.pop("shapes", True):
        shapes = tuple(tuple(op.shape) for op in shapes)
    kwargs.pop("einsum_call", True)
    if _real_oe is not None:  # pragma: no cover
        return _real_oe.contract_path(subscripts, *shapes, shapes=True, einsum_call=True, **kwargs)
    return shapes, _own_contract_path(subscripts, *shapes, **kwargs)


def path_flops(contraction_list, subscripts, shapes):
    """Path FLOPs in the opt_einsum / numpy convention."""
    terms, _ = _parse(subscripts, shapes)
    sizes = _size_dict(terms, shapes)
    total = 0
    for _, removed, step, _, _ in contraction_list:
        ins = step.split("->")[0].split(",")
        total += flop_count(set("".join(ins)), bool(removed), len(ins), sizes)
    return total


@lru_cache(maxsize=128)
def cached_contract_list(einstr, operand_shapes, optimize="auto", memory_limit=None, use_blas=True):
    """Memoised path lookup; mirrors ``_contract_path`` (``contractn/einsum.py:313-323``).

    As in the reference, an explicit path must be hashable (tuple of tuples).
    """
    _, lst = contract_path(
        einstr,
        *operand_shapes,
        shapes=True,
        einsum_call=True,
        optimize=optimize,
        memory_limit=memory_limit,
        use_blas=use_blas,
    )
    return lst
