"""The TN front end mirrors the reference graph model (no GPU needed).

Einsum strings are pinned against tests/golden/tn_strings.json, recorded from the
UNMODIFIED reference (oracle/make_tn_golden.py).  The structural checks follow
the reference's own tests (contractn/tests/test_ctn.py, test_nodes.py).
"""
import json
import os
from itertools import combinations
from math import prod
from string import ascii_lowercase as alph

import numpy as np
import pytest

from contractn_b200 import TN
from oracle.make_tn_golden_replay import replay

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tn_strings.json")
with open(GOLDEN) as f:
    CASES = json.load(f)


def index_inverse(tn):
    for node in tn.nodes():
        assert all(node.index(node[i]) == i for i in range(node.ndim))
        assert all(node[node.index(e)].name == e for e in node.edge_names)
        assert all(node[node.index(e)] == e for e in node.edges)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_einsum_string_matches_reference(case):
    tn = replay(TN, case["recipe"])
    assert tn.einsum_str == case["einsum_str"]
    index_inverse(tn)


def test_readme_strings():
    """README.md:56-57."""
    cp, tucker = TN(), TN()
    cp_hub = cp.add_copy_node(3)
    tucker_hub = tucker.add_dense_node(np.ones((4, 4, 4)))
    for i in range(3):
        mat = np.eye(4, 10)
        cp.connect_nodes(cp_hub, cp.add_dense_node(mat), i, 0)
        tucker.connect_nodes(tucker_hub, tucker.add_dense_node(mat), i, 0)
    assert cp.einsum_str == "ac,ad,ae->cde"
    assert tucker.einsum_str == "abc,ae,bf,cg->efg"


@pytest.mark.parametrize("node_type", ["dense", "hyper"])
@pytest.mark.parametrize("topology", ["path", "complete"])
@pytest.mark.parametrize("num_nodes", [2, 3, 6])
@pytest.mark.parametrize("neg", [False, True])
def test_connect_nodes(node_type, topology, num_nodes, neg):
    tn = TN()
    if topology == "path":
        make = (lambda: tn.add_dense_node(np.ones((3, 2, 2)))) if node_type == "dense" else (lambda: tn.add_copy_node(3, dim=2))
        nodes = [make() for _ in range(num_nodes)]
        for i in range(num_nodes - 1):
            tn.connect_nodes(nodes[i], nodes[i + 1], *((-1, -2) if neg else (2, 1)))
        assert len(tn.nodes()) == tn.num_cores == num_nodes
        assert tn.num_duplicate == tn.num_input == 0
        assert len(tn.nodes(danglers=True)) == 2 * num_nodes + 2
        assert len(tn.edges()) == 2 * num_nodes + 1
        n_sym = 2 * num_nodes + 1 if node_type == "dense" else 1
        assert len(tn.edge_symbols) == n_sym
        for i, node in enumerate(nodes):
            full = set(node.neighbors)
            near = set(nodes[max(i - 1, 0):i] + nodes[i + 1:i + 2])
            assert len(full - near) == 1 + int(i in (0, num_nodes - 1))
            assert near <= full
    else:
        if node_type == "dense":
            make = lambda: tn.add_dense_node(np.ones((3,) + (2,) * (num_nodes - 1)))
        else:
            make = lambda: tn.add_copy_node(num_nodes, dim=2)
        nodes = [make() for _ in range(num_nodes)]
        for i, j in combinations(range(num_nodes), 2):
            idx = (j - num_nodes, i + 1 - num_nodes) if neg else (j, i + 1)
            tn.connect_nodes(nodes[i], nodes[j], *idx)
        assert len(tn.nodes()) == tn.num_cores == num_nodes
        assert len(tn.edges()) == (num_nodes ** 2 + num_nodes) / 2
        assert len(tn.nodes(danglers=True)) == 2 * num_nodes
        n_sym = (num_nodes ** 2 + num_nodes) / 2 if node_type == "dense" else 1
        assert len(tn.edge_symbols) == n_sym
        for i, node in enumerate(nodes):
            full = set(node.neighbors)
            others = set(nodes[:i] + nodes[i + 1:])
            assert len(full) == len(others) + 1 and others <= full
    if node_type == "dense":
        assert tn.num_copy == 0 and tn.num_dense == num_nodes
    else:
        assert tn.num_dense == 0 and tn.num_copy == num_nodes
    before = tn.einsum_str
    tn._cleanup_edge_symbols()
    assert tn.einsum_str == before
    index_inverse(tn)


@pytest.mark.parametrize("num_nodes", [2, 4, 6])
@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("by_name", [False, True])
def test_remove_edges(num_nodes, single, by_name):
    tn = TN()
    nodes = [tn.add_dense_node(np.ones((2,) * (num_nodes - 1))) for _ in range(num_nodes)]
    for i, j in combinations(range(num_nodes), 2):
        tn.connect_nodes(nodes[i], nodes[j], j - 1, i)
    if single:
        for e in tn.edges():
            tn.remove_edge(e.name if by_name else e)
    else:
        tn.remove_edges_from([e.name for e in tn.edges()] if by_name else tn.edges())
    assert tn.num_dense == tn.num_cores == num_nodes
    assert tn.num_duplicate == tn.num_copy == tn.num_input == 0
    assert len(tn.edges()) == len(tn.edge_symbols) == num_nodes * (num_nodes - 1)
    assert all(e.dangler for e in tn.edges())


@pytest.mark.parametrize("order", [0, 1, 3])
@pytest.mark.parametrize("use_name", [False, True])
@pytest.mark.parametrize("use_symbols", [False, True])
def test_add_dense_and_duplicate_node(order, use_name, use_symbols):
    tn = TN()
    tensor = np.ones((2,) * order)
    name = "boring_core" if use_name else None
    syms = tuple(alph[26 - order:]) if use_symbols and order else None
    good = tuple(alph[:order]) if syms is None else syms
    node = tn.add_dense_node(tensor, name=name, edge_symbols=syms)
    assert node.node_type == "dense" and node.name == (name or "node_0")
    assert tn.num_dense == tn.num_cores == 1
    assert tn.edge_symbols == set(good)
    assert all(n.dangler for n in node.neighbors)
    assert node.ndim == order and node.shape == tensor.shape and node.edge_symbols == good
    assert node.size == node.numel == tensor.size
    for attr in ("base_node", "degree", "dim", "var_axes"):
        with pytest.raises(ValueError):
            getattr(node, attr)
    clone = tn.add_duplicate_node(node.name if use_name else node)
    assert clone.node_type == "clone" and clone.base_node is node
    assert tn.num_dense == tn.num_duplicate == 1 and tn.num_cores == 2
    assert clone.shape == tensor.shape and clone.size == tensor.size and clone.numel == 0
    for attr in ("tensor", "degree", "dim", "var_axes"):
        with pytest.raises(ValueError):
            getattr(clone, attr)
    index_inverse(tn)


@pytest.mark.parametrize("order", [0, 1, 3])
@pytest.mark.parametrize("use_dim", [False, True])
@pytest.mark.parametrize("symbols", [None, "z", "tuple"])
def test_add_copy_node(order, use_dim, symbols):
    tn = TN()
    dim = 5 if use_dim else None
    es = ("z",) * order if symbols == "tuple" else symbols
    if order == 0:
        with pytest.raises(Exception):
            tn.add_copy_node(order, dim=dim, edge_symbols=es)
        return
    node = tn.add_copy_node(order, dim=dim, edge_symbols=es)
    good = (("a" if es is None else "z"),) * order
    assert node.node_type == "hyper" and node.name == "node_0"
    assert tn.num_copy == tn.num_cores == 1
    assert tn.edge_symbols == set(good) and node.edge_symbols == good
    assert node.shape == ((5 if use_dim else -1),) * order
    assert node.size == (5 ** order if use_dim else None) and node.numel == 0
    for attr in ("tensor", "base_node", "var_axes"):
        with pytest.raises(ValueError):
            getattr(node, attr)
    index_inverse(tn)


@pytest.mark.parametrize("shape,var", [((), ()), ((4,), ()), ((3, 5, 2), (1,))])
def test_add_input_node(shape, var):
    tn = TN()
    node = tn.add_input_node(shape, var_shape_axes=var)
    assert node.node_type == "input" and tn.num_input == tn.num_cores == 1
    assert node.shape == tuple(-1 if i in var else d for i, d in enumerate(shape))
    assert node.size == (None if var else prod(shape)) and node.numel == 0
    assert node.edge_symbols == tuple(alph[:len(shape)])
    for attr in ("tensor", "base_node", "degree", "dim"):
        with pytest.raises(ValueError):
            getattr(node, attr)


def test_errors_mirror_reference():
    tn = TN()
    tn.add_dense_node(np.ones((2, 2)), name="x", edge_symbols="pq")
    with pytest.raises(TypeError, match="already in use"):
        tn.add_dense_node(np.ones((2,)), name="x")
    with pytest.raises(TypeError, match="already in use"):
        tn.add_dense_node(np.ones((2,)), edge_symbols="p")
    with pytest.raises(AssertionError):
        tn.add_dense_node([1, 2, 3])  # no .ndim / .shape
    a, b = tn.add_dense_node(np.ones((2, 3))), tn.add_dense_node(np.ones((4, 3)))
    with pytest.raises(AssertionError):
        tn.connect_nodes(a, b, 0, 0)  # 2 vs 4


def test_arg_packer_params_inputs_and_clones():
    tn = TN()
    a = tn.add_dense_node(np.ones((2, 3)))
    x = tn.add_input_node((3, 4))
    b = tn.add_duplicate_node(a)
    tn.connect_nodes(a, x, 1, 0)
    pack = tn._arg_packer()
    pa, px = np.zeros((2, 3)), np.ones((3, 4))
    ops = pack((pa,), (px,))
    assert ops[0] is pa and ops[1] is px and ops[2] is pa
    with pytest.raises(AssertionError):
        pack((pa, pa), (px,))
    assert tn.params == (a.tensor,)
