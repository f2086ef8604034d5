"""Tensor-network front end with ContracTN's user API, backed by the B200 engine.

This is the host-side mirror of the reference's graph model for users (and GPU
boxes) that do not have the reference package: same method names, argument
meaning, einsum-string output and error behaviour as

* ``TN``    -- /root/reference/contractn/ctn.py:11-486
* ``Node``  -- /root/reference/contractn/nodes.py:42-348
* ``Edge``  -- /root/reference/contractn/edges.py:4-50

but written from scratch on plain Python containers (ordered dicts and per-leg
edge records) instead of a NetworkX multigraph, with O(1) free-symbol
allocation instead of a scan over all edges per call.  Everything here is
bookkeeping: the arithmetic happens in ``contractn_b200.engine``.

Behaviour that is deliberately identical to the reference, because it decides
the einsum string (and so the contraction path):

* a node's legs get the lowest unused symbol indices (gap filling,
  ``contractn/utils.py:44-63``), a copy node gets ONE symbol for all legs;
* ``connect_nodes`` labels the new bond ``min(sym1, sym2)`` by code point
  (``ctn.py:199``) and frees the other symbol;
* dangling legs are explicit placeholder nodes, listed in creation order; the
  output subscript is the dangling symbols in that order (``einsum.py:124-130``);
* after an edge touches a copy node, the dangling legs of the connected cluster
  of copy nodes are relabelled to the cluster's most common symbol
  (``ctn.py:271-313``) -- which is why copy nodes must be created BEFORE the
  dense nodes they attach to (SURVEY.md finding 0.7).
"""
from __future__ import annotations

import heapq
from collections import Counter
from math import prod

from .planner import get_symbol

_LETTERS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
_FULL_NAMES = {"dense": "dense", "clone": "duplicate", "hyper": "copy", "input": "input"}
_NODE_KWARGS = {
    "dense": ({"tensor"}, set()),
    "clone": ({"base_node"}, set()),
    "hyper": ({"degree"}, {"dim"}),
    "input": ({"shape", "var_axes"}, set()),
    "dangler": (set(), set()),
}


def _symbol_index(symbol):
    """Inverse of ``get_symbol`` (``contractn/utils.py:67-81``)."""
    assert isinstance(symbol, str) and len(symbol) == 1
    pos = _LETTERS.find(symbol)
    idx = pos if pos >= 0 else ord(symbol) - 140
    assert idx >= 0 and get_symbol(idx) == symbol
    return idx


def _check_symbol(symbol):
    assert isinstance(symbol, str)
    assert len(symbol) == 1


def _wrong_type(node_type, attr, node_name, actual):
    return ValueError(
        f"Only {_FULL_NAMES[node_type]} nodes have {attr} attributes "
        f"(node '{node_name}' has node type '{actual}')"
    )


class Edge:
    """One bond (or dangling leg) of the network."""

    def __init__(self, tn, name, dim, symbol, ends):
        assert isinstance(dim, int)
        _check_symbol(symbol)
        self.tn = tn
        self.name = name  # (node name, node name, parallel-edge key)
        self._dim = dim
        self._symbol = symbol
        self._ends = ends  # ((Node, leg), (Node, leg))

    @property
    def symbol(self):
        return self._symbol

    @property
    def dim(self):
        return self._dim

    @property
    def var_dim(self):
        return self._dim < 0

    @property
    def nodes(self):
        return tuple(n for n, _ in self._ends)

    @property
    def dangler(self):
        return any(n.dangler for n in self.nodes)

    def __repr__(self):
        return f"Edge(name={self.name}, symbol={self._symbol!r}, dim={self._dim})"


class Node:
    """A core tensor, copy tensor, input slot or dangling-leg placeholder."""

    def __init__(self, tn, node_type, name, n_legs, **kwargs):
        required, optional = _NODE_KWARGS[node_type]
        given = set(kwargs)
        if not given >= required:
            missing = (required - given).pop()
            raise TypeError(f"Argument '{missing}' missing, needed for node_type '{node_type}'")
        if not (required | optional) >= given:
            extra = (given - required - optional).pop()
            raise TypeError(f"Argument '{extra}' not recognized for node_type '{node_type}'")
        self.tn = tn
        self.name = name
        self.node_type = node_type
        self._legs = [None] * n_legs  # Edge per leg
        self._attrs = kwargs
        if node_type == "dense":
            assert n_legs == kwargs["tensor"].ndim
        elif node_type == "clone":
            base = kwargs["base_node"]
            assert base.node_type == "dense"
            assert n_legs == base.tensor.ndim
        elif node_type == "hyper":
            assert kwargs["degree"] > 0, "Hyperedge nodes must have positive degree"
            assert n_legs == kwargs["degree"]
            kwargs.setdefault("dim", None)
            assert isinstance(kwargs["dim"], int) or kwargs["dim"] is None
        elif node_type == "input":
            axes = kwargs["var_axes"]
            assert n_legs == len(kwargs["shape"])
            assert len(set(axes)) == len(axes)
            assert all(0 <= a < n_legs for a in axes)

    # -- classification ----------------------------------------------------
    @property
    def dangler(self):
        return self.node_type == "dangler"

    @property
    def is_copy(self):
        return self.node_type == "hyper"

    # -- shape -------------------------------------------------------------
    @property
    def shape(self):
        """Tensor shape; -1 marks a not-yet-determined extent."""
        kind = self.node_type
        if kind == "dense":
            return tuple(self._attrs["tensor"].shape)
        if kind == "clone":
            return tuple(self._attrs["base_node"].tensor.shape)
        if kind == "hyper":
            d = self._attrs["dim"]
            return (-1 if d is None else d,) * self._attrs["degree"]
        if kind == "input":
            var = self._attrs["var_axes"]
            return tuple(-1 if i in var else d for i, d in enumerate(self._attrs["shape"]))
        return (-1,)

    @property
    def ndim(self):
        return len(self._legs)

    @property
    def size(self):
        shp = self.shape
        return None if any(d < 0 for d in shp) else prod(shp)

    @property
    def numel(self):
        return prod(self.tensor.shape) if self.node_type == "dense" else 0

    # -- type specific attributes -----------------------------------------
    def _typed(self, node_type, attr, key):
        if self.node_type != node_type:
            raise _wrong_type(node_type, attr, self.name, self.node_type)
        return self._attrs[key]

    @property
    def tensor(self):
        return self._typed("dense", "tensor", "tensor")

    @tensor.setter
    def tensor(self, array):
        self._typed("dense", "tensor", "tensor")
        assert hasattr(array, "ndim") and hasattr(array, "shape")
        assert array.ndim == self.ndim
        self._attrs["tensor"] = array

    @property
    def base_node(self):
        return self._typed("clone", "base_node", "base_node")

    @property
    def degree(self):
        return self._typed("hyper", "degree", "degree")

    @property
    def dim(self):
        return self._typed("hyper", "dim", "dim")

    @property
    def var_axes(self):
        return self._typed("input", "var_axes", "var_axes")

    # -- connectivity ------------------------------------------------------
    @property
    def edges(self):
        return tuple(self._legs)

    @property
    def edge_names(self):
        return [e.name for e in self._legs]

    @property
    def edge_symbols(self):
        return tuple(e.symbol for e in self._legs)

    @property
    def symbol(self):
        assert self.dangler
        return self._legs[0].symbol

    def _other_end(self, leg):
        end0, end1 = self._legs[leg]._ends
        return end1[0] if (end0[0] is self and end0[1] == leg) else end0[0]

    @property
    def neighbors(self):
        """Node at the far end of each leg, in leg order (stubs for open legs)."""
        return tuple(self._other_end(leg) for leg in range(len(self._legs)))

    def __getitem__(self, leg):
        return self._legs[leg]

    def index(self, edge):
        if isinstance(edge, Edge):
            edge = edge.name
        names = self.edge_names
        assert edge in names
        return names.index(edge)

    def __repr__(self):
        return f"Node(name={self.name}, node_type={self.node_type}, degree={self.ndim})"


class TN:
    """Tensor network with copy nodes; contraction runs on the B200 engine."""

    def __init__(self):
        self._nodes = {}  # name -> Node, insertion ordered, includes danglers
        self._edges = {}  # edge name -> Edge
        self._dang_id = 0
        self._n_cores = 0
        self._sym_refs = Counter()  # symbol index -> number of edges carrying it
        self._free = []  # heap of freed symbol indices below _next_sym
        self._next_sym = 0
        self._pair_keys = {}  # frozenset({n1, n2}) -> set of used parallel keys
        self._version = 0  # bumped on every structural change (plan cache key)
        self._compiled = None

    # ------------------------------------------------------------------
    # symbols
    # ------------------------------------------------------------------
    def _take_symbols(self, count):
        """The ``count`` lowest unused symbol indices (``utils.py:44-63``)."""
        taken = []
        while len(taken) < count:
            if self._free:
                idx = heapq.heappop(self._free)
                if self._sym_refs[idx] > 0 or idx in taken:
                    continue
            else:
                idx = self._next_sym
                self._next_sym += 1
                if self._sym_refs[idx] > 0:
                    continue
            taken.append(idx)
        # not referenced yet: hand them back so an aborted call leaks nothing
        for idx in taken:
            heapq.heappush(self._free, idx)
        return tuple(get_symbol(i) for i in taken)

    def _ref(self, symbol):
        idx = _symbol_index(symbol)
        self._sym_refs[idx] += 1
        if idx >= self._next_sym:
            for gap in range(self._next_sym, idx):
                heapq.heappush(self._free, gap)
            self._next_sym = idx + 1

    def _unref(self, symbol):
        idx = _symbol_index(symbol)
        self._sym_refs[idx] -= 1
        if self._sym_refs[idx] == 0:
            del self._sym_refs[idx]
            heapq.heappush(self._free, idx)

    def _new_edge_symbols(self, node_type, degree, edge_symbols=None):
        if edge_symbols is not None:
            assert len(edge_symbols) == degree
            for s in edge_symbols:
                _check_symbol(s)
            clash = self.edge_symbols.intersection(edge_symbols)
            if clash:
                raise TypeError(f"Edge symbol '{clash.pop()}' already in use")
            return tuple(edge_symbols)
        if degree == 0:
            return tuple()
        assert degree > 0
        assert node_type != "dangler"
        if node_type == "hyper":
            return self._take_symbols(1) * degree
        return self._take_symbols(degree)

    # ------------------------------------------------------------------
    # node / edge creation
    # ------------------------------------------------------------------
    def _new_node_name(self, name=None):
        if name is None:
            name = f"node_{self.num_cores}"
        assert isinstance(name, str)
        if name in self._nodes:
            raise TypeError(f"Node name '{name}' already in use")
        return name

    def _make_edge(self, node1, leg1, node2, leg2, symbol):
        assert isinstance(node1, Node) and isinstance(node2, Node)
        _check_symbol(symbol)
        assert -node1.ndim <= leg1 < node1.ndim
        assert -node2.ndim <= leg2 < node2.ndim
        assert node1 in self and node2 in self
        leg1 %= node1.ndim
        leg2 %= node2.ndim
        d1, d2 = node1.shape[leg1], node2.shape[leg2]
        if d1 < 0 and d2 < 0:
            dim = -1
        elif (d1 < 0) != (d2 < 0):
            dim = max(d1, d2)
        else:
            assert d1 == d2
            dim = d1
        pair = frozenset((node1.name, node2.name))
        used = self._pair_keys.setdefault(pair, set())
        key = len(used)
        while key in used:
            key += 1
        used.add(key)
        edge = Edge(self, (node1.name, node2.name, key), int(dim), symbol,
                    ((node1, leg1), (node2, leg2)))
        self._edges[edge.name] = edge
        self._ref(symbol)
        node1._legs[leg1] = edge
        node2._legs[leg2] = edge
        if "hyper" in (node1.node_type, node2.node_type):
            self._cleanup_edge_symbols(node1 if node1.is_copy else node2)
        return edge

    def _drop_edge(self, edge):
        del self._edges[edge.name]
        self._pair_keys[frozenset(edge.name[:2])].discard(edge.name[2])
        self._unref(edge.symbol)

    def _new_dangler(self, parent, leg, symbol):
        name = f"_dangler_{self._dang_id}"
        assert name not in self._nodes and parent.name in self._nodes
        self._dang_id += 1
        stub = Node(self, "dangler", name, 1)
        self._nodes[name] = stub
        return self._make_edge(parent, leg, stub, 0, symbol)

    def _init_node(self, node_type, name, edge_symbols, **kwargs):
        name = self._new_node_name(name)
        assert node_type != "dangler"
        if node_type == "hyper":
            assert len(set(edge_symbols)) == 1
        else:
            assert len(set(edge_symbols)) == len(edge_symbols)
        node = Node(self, node_type, name, len(edge_symbols), **kwargs)
        self._nodes[name] = node
        self._n_cores += 1
        for leg, s in enumerate(edge_symbols):
            self._new_dangler(node, leg, s)
        self._version += 1
        return node

    def add_dense_node(self, tensor, name=None, edge_symbols=None):
        """Add a dense core holding ``tensor`` (``ctn.py:130``)."""
        assert hasattr(tensor, "ndim")
        assert hasattr(tensor, "shape")
        syms = self._new_edge_symbols("dense", tensor.ndim, edge_symbols)
        return self._init_node("dense", name, syms, tensor=tensor)

    def add_duplicate_node(self, base_node, name=None, edge_symbols=None):
        """Add a weight-sharing clone of a dense node (``ctn.py:141``)."""
        if not isinstance(base_node, Node):
            assert base_node in self._nodes
            base_node = self._nodes[base_node]
        syms = self._new_edge_symbols("clone", base_node.ndim, edge_symbols)
        return self._init_node("clone", name, syms, base_node=base_node)

    def add_copy_node(self, degree, dim=None, name=None, edge_symbols=None):
        """Add a copy tensor (hyperedge) of the given order (``ctn.py:154``)."""
        if isinstance(edge_symbols, str):
            edge_symbols = [edge_symbols] * degree
        if edge_symbols is not None:
            assert len(edge_symbols) == degree
            for s in edge_symbols:
                _check_symbol(s)
            clash = self.edge_symbols.intersection(edge_symbols)
            if clash:
                raise TypeError(f"Edge symbol '{clash.pop()}' already in use")
            syms = tuple(edge_symbols)
        else:
            syms = self._new_edge_symbols("hyper", degree)
        kwargs = {"degree": degree}
        if dim is not None:
            kwargs["dim"] = dim
        return self._init_node("hyper", name, syms, **kwargs)

    def add_input_node(self, shape, var_shape_axes=(), name=None, edge_symbols=None):
        """Add a placeholder filled at contraction time (``ctn.py:167``)."""
        syms = self._new_edge_symbols("input", len(shape), edge_symbols)
        return self._init_node("input", name, syms, shape=shape, var_axes=var_shape_axes)

    # ------------------------------------------------------------------
    # wiring
    # ------------------------------------------------------------------
    def connect_nodes(self, node1, node2, index1, index2, edge_symbol=None):
        """Join two dangling legs into a bond (``ctn.py:179-204``)."""
        if not isinstance(node1, Node):
            assert node1 in self._nodes
            node1 = self._nodes[node1]
        if not isinstance(node2, Node):
            assert node2 in self._nodes
            node2 = self._nodes[node2]
        old1, old2 = node1[index1], node2[index2]
        s1, s2 = old1.symbol, old2.symbol
        if edge_symbol in self.edge_symbols:
            assert edge_symbol in {s1, s2}
        if edge_symbol is None:
            edge_symbol = min(s1, s2)
        for owner, leg, old in ((node1, index1, old1), (node2, index2, old2)):
            stub = owner._other_end(leg % owner.ndim)
            assert stub.dangler, "only dangling legs can be connected"
            del self._nodes[stub.name]
            self._drop_edge(old)
        self._make_edge(node1, index1, node2, index2, edge_symbol)
        self._version += 1

    def _remove_edges(self, edge_set):
        assert all(isinstance(e, Edge) for e in edge_set)
        assert all(e.name in self._edges for e in edge_set)
        hit = []
        for e in edge_set:
            if e.dangler:
                continue
            for node, leg in e._ends:
                hit.append((node, leg))
        for e in edge_set:
            if not e.dangler:
                self._drop_edge(e)
        fresh = self._take_symbols(len(hit))
        for (node, leg), s in zip(hit, fresh):
            self._new_dangler(node, leg, s)
        if any(node.is_copy for node, _ in hit):
            self._cleanup_edge_symbols()
        self._version += 1

    def remove_edge(self, edge):
        """Cut one bond, leaving two dangling legs (``ctn.py:206``)."""
        assert isinstance(edge, (Edge, tuple))
        if isinstance(edge, tuple):
            edge = self._edges[edge]
        self._remove_edges({edge})

    def remove_edges_from(self, edge_set):
        """Cut several bonds; already-dangling legs are left alone (``ctn.py:215``)."""
        edge_set = tuple(edge_set)
        assert all(isinstance(e, (Edge, tuple)) for e in edge_set)
        self._remove_edges({self._edges[e] if isinstance(e, tuple) else e for e in edge_set})

    def _relabel(self, edge, symbol):
        if edge.symbol != symbol:
            self._unref(edge.symbol)
            edge._symbol = symbol
            self._ref(symbol)

    def _copy_cluster(self, seed):
        """Connected component of ``seed`` among copy nodes and dangling stubs."""
        seen, stack = {seed.name: seed}, [seed]
        while stack:
            node = stack.pop()
            for leg, e in enumerate(node._legs):
                if e is None:
                    continue  # leg not wired yet (node under construction)
                other = node._other_end(leg)
                if (other.is_copy or other.dangler) and other.name not in seen:
                    seen[other.name] = other
                    stack.append(other)
        return seen

    def _cleanup_edge_symbols(self, naughty_node=None):
        """Give each cluster of connected copy nodes one symbol (``ctn.py:271-313``)."""
        if naughty_node is not None:
            assert naughty_node.is_copy
            clusters = [self._copy_cluster(naughty_node)]
        else:
            clusters, done = [], set()
            for node in self._nodes.values():
                if (node.is_copy or node.dangler) and node.name not in done:
                    comp = self._copy_cluster(node)
                    done.update(comp)
                    clusters.append(comp)
        for comp in clusters:
            votes, seen = Counter(), set()
            for node in comp.values():
                for e in node._legs:
                    if e is not None and id(e) not in seen:
                        seen.add(id(e))
                        votes[e.symbol] += 1
            best = votes.most_common(1)[0][0]
            _check_symbol(best)
            for node in comp.values():
                for e in node._legs:
                    if e is not None and all(n.name in comp for n in e.nodes):
                        self._relabel(e, best)

    # ------------------------------------------------------------------
    # views
    # ------------------------------------------------------------------
    def nodes(self, as_iter=False, copy_nodes=True, danglers=False):
        """Nodes in insertion order (``ctn.py:315-336``)."""
        it = (
            n for n in self._nodes.values()
            if (copy_nodes or not n.is_copy) and (danglers or not n.dangler)
        )
        return it if as_iter else tuple(it)

    def edges(self, as_iter=False):
        it = iter(self._edges.values())
        return it if as_iter else tuple(it)

    @property
    def edge_symbols(self):
        return {get_symbol(i) for i, c in self._sym_refs.items() if c > 0}

    @property
    def einsum_str(self):
        """einsum-style string of the whole network (``einsum.py:117-130``)."""
        terms, free = [], []
        for node in self._nodes.values():
            if node.is_copy:
                continue
            if node.dangler:
                free.append(node.symbol)
            else:
                terms.append("".join(node.edge_symbols))
        return ",".join(terms) + "->" + "".join(free)

    @property
    def params(self):
        """Unique parameter tensors, in node insertion order (``ctn.py:418-428``)."""
        return tuple(n.tensor for n in self._nodes.values() if n.node_type == "dense")

    def _count(self, node_type):
        return sum(1 for n in self._nodes.values() if n.node_type == node_type)

    @property
    def num_dense(self):
        return self._count("dense")

    @property
    def num_duplicate(self):
        return self._count("clone")

    @property
    def num_copy(self):
        return self._count("hyper")

    @property
    def num_input(self):
        return self._count("input")

    @property
    def num_cores(self):
        return self._n_cores

    def __contains__(self, node):
        if isinstance(node, Node):
            node = node.name
        return node in self._nodes

    def __getitem__(self, name):
        assert name in self._nodes
        return self._nodes[name]

    # ------------------------------------------------------------------
    # contraction
    # ------------------------------------------------------------------
    def _arg_packer(self):
        """(params, inputs) -> einsum operand tuple (``einsum.py:133-187``).

        Clone nodes are wired to their base node's parameter slot (the reference
        tests ``"duplicate"`` where nodes are typed ``"clone"`` and so asserts,
        SURVEY.md section 8f item 2; here weight sharing works).
        """
        param_slots, input_slots, dense_pos, count = [], [], {}, 0
        for node in self.nodes(as_iter=True, copy_nodes=False, danglers=False):
            if node.node_type == "dense":
                dense_pos[node.name] = len(param_slots)
                param_slots.append([count])
            elif node.node_type == "clone":
                base = node.base_node.name
                assert base in dense_pos
                param_slots[dense_pos[base]].append(count)
            elif node.node_type == "input":
                input_slots.append(count)
            count += 1
        n_par, n_in = len(param_slots), len(input_slots)

        def arg_packer(params, inputs):
            assert len(params) == n_par
            assert len(inputs) == n_in
            ops = [None] * count
            for p, slots in zip(params, param_slots):
                for s in slots:
                    ops[s] = p
            for x, s in zip(inputs, input_slots):
                ops[s] = x
            assert all(op is not None for op in ops)
            return tuple(ops)

        return arg_packer

    def make_contract_fun(self, optimize="auto", split_format=False):
        """``contract_fun(params, inputs)`` closure (``ctn.py:349-387``).

        The string and argument packer are compiled once per network version
        (the reference recompiles on every ``TN.contract`` call, SURVEY.md
        section 8f item 1).
        """
        from .engine import contract

        if self._compiled is None or self._compiled[0] != self._version:
            self._compiled = (self._version, self.einsum_str, self._arg_packer())
        _, einstr, packer = self._compiled

        def contract_fun(params, inputs):
            operands = packer(params, inputs)
            return contract(einstr, *operands, optimize=optimize, split_format=split_format)

        return contract_fun

    def contract(self, inputs=(), optimize="auto", split_format=False):
        """Contract the network to one dense tensor (``ctn.py:389-409``)."""
        fun = self.make_contract_fun(optimize=optimize, split_format=split_format)
        return fun(self.params, inputs)
