"""Builders for the BASELINE.json networks and explicit sweep paths for them.

The reference leaves path choice to ``opt_einsum``; its greedy optimiser is poor on
grids and on an MPS with a batch hub, so the
benchmarks and tests of cfg 3 and cfg 5 pass an explicit path -- a tuple of tuples, the
only explicit form ``contract`` accepts (``contractn/einsum.py:313`` hashes it).  The
generators here produce those paths in the reference's linear format from an SSA
description of the sweep.

Every builder follows the reference's rule that copy nodes are created BEFORE the
dense nodes they attach to (SURVEY.md finding 0.7, ``contractn/ctn.py:199``).
"""
from __future__ import annotations

import numpy as np

from .planner import _ssa_to_linear
from .tn import TN

__all__ = ["chain_tn", "hub_tn", "mps_tn", "cp_tn", "tucker_tn", "peps_tn",
           "mps_sweep_path", "grid_sweep_path", "ssa_to_linear"]


def ssa_to_linear(ssa_path, n_inputs):
    """SSA ids (every intermediate gets a fresh id) -> the positional ("linear") path
    format ``contract(optimize=...)`` takes."""
    return tuple(_ssa_to_linear(list(ssa_path), n_inputs))


def mps_sweep_path(sites):
    """Left-to-right sweep of a batched MPS inner product whose operands are ordered
    cores ``0..sites-1`` then inputs ``sites..2*sites-1`` (the order ``mps_tn`` creates
    them in): boundary x core, then x input, site by site -- every step is the GEMM
    ``br,prs->bps`` followed by the hyperedge step ``bps,bp->bs`` (cfg 3)."""
    ssa, nxt = [(0, sites)], 2 * sites
    cur = nxt
    nxt += 1
    for i in range(1, sites - 1):
        ssa.append((cur, i))
        t = nxt
        nxt += 1
        ssa.append((t, sites + i))
        cur = nxt
        nxt += 1
    ssa.append((sites - 1, 2 * sites - 1))
    v = nxt
    nxt += 1
    ssa.append((cur, v))
    return ssa_to_linear(ssa, 2 * sites)


def grid_sweep_path(n_nodes):
    """Row-by-row boundary sweep of a grid whose nodes were created in row-major order
    (cfg 5): the boundary absorbs one site per step."""
    return tuple([(0, 1)] + [(0, n_nodes - t - 1) for t in range(1, n_nodes - 1)])


def chain_tn(n=1000, random=False, seed=1001):
    """cfg 1, README verbatim (``README.md:66-71``): a vector and ``n`` 3x3 matrices."""
    rng = np.random.default_rng(seed)
    tn = TN()
    prev = tn.add_dense_node(rng.uniform(0.5, 1.5, (3,)) if random else np.ones((3,)))
    for _ in range(n):
        m = tn.add_dense_node(rng.uniform(0.5, 1.5, (3, 3)) if random else np.ones((3, 3)))
        tn.connect_nodes(prev, m, -1, 0)
        prev = m
    return tn, "auto"


def hub_tn(n=100, batch=1 << 20, dtype=np.float64, seed=1002, arrays=None):
    """cfg 2: hub ``a`` first, batch hub ``b``, then ``n`` dense (2, batch) nodes."""
    rng = np.random.default_rng(seed)
    tn = TN()
    hub, hubb = tn.add_copy_node(n + 1), tn.add_copy_node(n + 1)
    for i in range(n):
        arr = arrays[i] if arrays is not None else rng.uniform(0.9, 1.1, (2, batch)).astype(dtype)
        v = tn.add_dense_node(arr)
        tn.connect_nodes(hub, v, i, 0)
        tn.connect_nodes(hubb, v, i, 1)
    return tn, "auto"


def mps_tn(sites=512, bond=512, phys=2, batch=4096, dtype=np.float32, seed=1003):
    """cfg 3: batch hub first, cores, then inputs; explicit left-to-right sweep path."""
    rng = np.random.default_rng(seed)
    tn = TN()
    hubb = tn.add_copy_node(sites + 1)
    cores = []
    for i in range(sites):
        shape = (phys, bond) if i in (0, sites - 1) else (phys, bond, bond)
        cores.append(tn.add_dense_node(rng.uniform(0, 1, shape).astype(dtype)))
    for i in range(sites - 1):
        tn.connect_nodes(cores[i], cores[i + 1], -1, 1)
    for i in range(sites):
        x = tn.add_dense_node(rng.uniform(0.9, 1.1, (batch, phys)).astype(dtype))
        tn.connect_nodes(cores[i], x, 0, 1)
        tn.connect_nodes(hubb, x, i, 0)
    return tn, mps_sweep_path(sites)


def cp_tn(I=1024, R=1024, dtype=np.float64, seed=1004):
    """cfg 4, CP reconstruction: hub first, three (I, R) factors on its legs."""
    rng = np.random.default_rng(seed)
    tn = TN()
    hub = tn.add_copy_node(3)
    for i in range(3):
        m = tn.add_dense_node(rng.uniform(0, 1, (I, R)).astype(dtype))
        tn.connect_nodes(hub, m, i, 1)
    return tn, "auto"


def tucker_tn(I=1024, R=256, dtype=np.float64, seed=1004):
    """cfg 4, Tucker reconstruction: (R, R, R) core and three (I, R) factors."""
    rng = np.random.default_rng(seed)
    tn = TN()
    core = tn.add_dense_node(rng.uniform(0, 1, (R, R, R)).astype(dtype))
    for i in range(3):
        m = tn.add_dense_node(rng.uniform(0, 1, (I, R)).astype(dtype))
        tn.connect_nodes(core, m, i, 1)
    return tn, "auto"


def peps_tn(L=8, D=8, dtype=np.float64, seed=1005):
    """cfg 5: closed LxL grid, leg order (up, left, down, right); row-by-row boundary sweep."""
    rng = np.random.default_rng(seed)
    tn = TN()
    grid = {}
    for r in range(L):
        for c in range(L):
            legs = [d for d, ok in (("u", r > 0), ("l", c > 0), ("d", r < L - 1), ("r", c < L - 1)) if ok]
            grid[r, c] = (tn.add_dense_node(rng.uniform(0, 1, (D,) * len(legs)).astype(dtype)), legs)
    for r in range(L):
        for c in range(L):
            node, legs = grid[r, c]
            if c < L - 1:
                o, ol = grid[r, c + 1]
                tn.connect_nodes(node, o, legs.index("r"), ol.index("l"))
            if r < L - 1:
                o, ol = grid[r + 1, c]
                tn.connect_nodes(node, o, legs.index("d"), ol.index("u"))
    return tn, grid_sweep_path(L * L)
