"""Host planner (opt_einsum 3.3.0 restatement): list format, classifier,
determinism and path quality.  No GPU needed."""
import subprocess
import sys

import numpy as np
import pytest

from contractn_b200 import planner

from _cases import Golden, golden_names


def test_symbols_round_trip():
    from contractn_b200.tn import _symbol_index

    assert planner.get_symbol(0) == "a" and planner.get_symbol(25) == "z"
    assert planner.get_symbol(26) == "A" and planner.get_symbol(51) == "Z"
    assert planner.get_symbol(52) == chr(192)
    for i in (0, 7, 51, 52, 53, 400, 1500):
        assert _symbol_index(planner.get_symbol(i)) == i


@pytest.mark.parametrize(
    "inputs,result,removed,want",
    [
        (["ij", "jk"], "ik", "j", "GEMM"),
        (["ijj", "jk"], "ik", "j", False),
        (["ab", "cd"], "abcd", "", "OUTER/EINSUM"),
        (["ab", "ab"], "", "ab", "DOT"),
        (["ab", "ba"], "", "ab", "DOT/EINSUM"),
        (["ab", "b"], "a", "b", "GEMM"),
        (["br", "prs"], "bps", "r", "TDOT"),
        (["dbc", "eb"], "dce", "b", "TDOT"),
        (["ab", "ab"], "ab", "", False),       # hyperedge / batch index
        (["ca", "ba"], "cab", "", False),      # Khatri-Rao
        (["bps", "bp"], "bs", "p", False),     # batch index b in both and out
        (["ab", "bc"], "c", "b", False),       # 'a' summed out of one operand
        (["abc", "bca"], "", "abc", "DOT/EINSUM"),
        (["abcd", "cd"], "ab", "cd", "GEMM"),
        (["cdab", "cd"], "ab", "cd", "GEMM"),
        (["acbd", "cd"], "ab", "cd", "GEMV/EINSUM"),
        (["ab", "bc", "cd"], "ad", "bc", False),
    ],
)
def test_can_blas(inputs, result, removed, want):
    assert planner.can_blas(inputs, set(result), set(removed)) == want


def test_can_blas_broadcast_is_not_blas():
    assert planner.can_blas(["ij", "jk"], set("ik"), set("j"), [(3, 1), (4, 5)]) is False


def test_list_format():
    shapes = [(3, 4), (4, 5), (5, 6)]
    args, lst = planner.contract_path("ab,bc,cd->ad", *shapes, shapes=True, einsum_call=True)
    assert len(lst) == 2
    for positions, removed, step, remaining, flag in lst:
        assert list(positions) == sorted(positions, reverse=True)
        assert isinstance(removed, (set, frozenset))
        ins, out = step.split("->")
        assert len(ins.split(",")) == len(positions)
        assert remaining is None or isinstance(remaining, tuple)
    assert lst[-1][2].endswith("->ad")  # last step writes the requested output order
    assert lst[-1][3] == ("ad",)
    # more than 20 operands left: `remaining` is dropped
    lst = planner.contract_path(",".join(["ab"] * 30) + "->ab", *[(2, 2)] * 30)[1]
    assert lst[0][3] is None and lst[-1][3] == ("ab",)


def test_auto_dispatch_matches_named_optimizers():
    rng = np.random.default_rng(1)

    def chain(n):
        terms = [planner.get_symbol(i) + planner.get_symbol(i + 1) for i in range(n)]
        shapes = [(int(rng.integers(2, 6)), 1) for _ in range(n)]
        dims = [int(rng.integers(2, 7)) for _ in range(n + 1)]
        shapes = [(dims[i], dims[i + 1]) for i in range(n)]
        return ",".join(terms) + "->" + planner.get_symbol(0) + planner.get_symbol(n), shapes

    def key(lst):
        return [(c[0], c[2]) for c in lst]

    for n, name in [(3, "optimal"), (4, "optimal"), (5, "branch-all"), (6, "branch-all"), (7, "branch-2"),
                    (8, "branch-2"), (9, "branch-1"), (14, "branch-1"), (15, "greedy"), (40, "greedy")]:
        s, shapes = chain(n)
        a = planner.contract_path(s, *shapes, optimize="auto")[1]
        b = planner.contract_path(s, *shapes, optimize=name)[1]
        assert key(a) == key(b), (n, name)


def test_optimal_is_no_worse_than_heuristics():
    rng = np.random.default_rng(2)
    for trial in range(25):
        n = int(rng.integers(3, 6))
        syms = "abcdefg"
        terms, shapes = [], []
        dims = {s: int(rng.integers(2, 6)) for s in syms}
        for _ in range(n):
            k = int(rng.integers(1, 4))
            t = "".join(rng.choice(list(syms), size=k, replace=False))
            terms.append(t)
            shapes.append(tuple(dims[c] for c in t))
        flat = "".join(terms)
        out = "".join(sorted(c for c in set(flat) if flat.count(c) == 1))
        s = ",".join(terms) + "->" + out
        cost = {}
        for opt in ("optimal", "branch-all", "branch-2", "branch-1", "greedy"):
            lst = planner.contract_path(s, *shapes, optimize=opt)[1]
            cost[opt] = planner.path_flops(lst, s, shapes)
        assert cost["optimal"] <= min(cost.values()), (s, cost)


def test_explicit_path_and_errors():
    shapes = [(3, 4), (4, 5), (5, 6)]
    lst = planner.contract_path("ab,bc,cd->ad", *shapes, optimize=((1, 2), (0, 1)))[1]
    assert lst[0][0] == (2, 1) and lst[0][2] == "cd,bc->db"
    with pytest.raises(KeyError):
        planner.contract_path("ab,bc,cd->ad", *shapes, optimize="no-such-optimizer")
    with pytest.raises(ValueError):
        planner.contract_path("ab,bc->ac", (3, 4), (5, 6))
    with pytest.raises(ValueError):
        planner.contract_path("ab,bc->ac", (3, 4))
    with pytest.raises(TypeError):
        planner.contract_path("ab,bc->ac", (3, 4), (4, 5), bogus=1)
    # like the reference (lru_cache), an explicit path must be hashable
    with pytest.raises(TypeError):
        planner.cached_contract_list("ab,bc,cd->ad", tuple(shapes), optimize=[(1, 2), (0, 1)])


def test_memory_limit_forces_nary_step():
    shapes = [(3, 4), (4, 5), (5, 6), (6, 2)]
    lst = planner.contract_path("ab,bc,cd,de->ae", *shapes, optimize="optimal", memory_limit=5)[1]
    assert len(lst) == 1 and len(lst[0][0]) == 4 and lst[0][4] is False
    lst = planner.contract_path("ab,bc,cd,de->ae", *shapes, optimize="greedy", memory_limit="max_input")[1]
    assert sum(len(c[0]) - 1 for c in lst) == 3


def test_greedy_folds_identical_terms_left_to_right():
    """cfg 2: the hub is 99 Hadamard steps ((v0*v1)*v2)*...."""
    lst = planner.contract_path(",".join(["ab"] * 100) + "->ab", *[(2, 8)] * 100)[1]
    assert len(lst) == 99
    assert all(c[2] == "ab,ab->ab" and c[4] is False for c in lst)
    assert lst[0][0] == (1, 0)


@pytest.mark.parametrize("name", golden_names())
def test_planner_reproduces_recorded_lists(name):
    g = Golden(name)
    shapes = tuple(tuple(o.shape) for o in g.operands)
    lst = planner.contract_path(g.einstr, *shapes, optimize=g.optimize, memory_limit=g.memory_limit)[1]
    assert [(tuple(c[0]), set(c[1]), c[2], c[4]) for c in lst] == \
           [(tuple(c[0]), set(c[1]), c[2], c[4]) for c in g.contraction_list]


def test_paths_do_not_depend_on_hash_seed():
    """Set iteration order must not leak into the path (PYTHONHASHSEED)."""
    code = (
        "from contractn_b200 import planner\n"
        "import numpy as np\n"
        "rng=np.random.default_rng(5)\n"
        "L=4\n"
        "import itertools\n"
        "terms=[];shapes=[];k=0;sym={}\n"
        "def s(key):\n"
        "    global k\n"
        "    if key not in sym:\n"
        "        sym[key]=planner.get_symbol(k);k+=1\n"
        "    return sym[key]\n"
        "for r in range(L):\n"
        "    for c in range(L):\n"
        "        t=''\n"
        "        if r>0:t+=s(('v',r-1,c))\n"
        "        if c>0:t+=s(('h',r,c-1))\n"
        "        if r<L-1:t+=s(('v',r,c))\n"
        "        if c<L-1:t+=s(('h',r,c))\n"
        "        terms.append(t);shapes.append((3,)*len(t))\n"
        "for opt in ('greedy','branch-1'):\n"
        "    lst=planner.contract_path(','.join(terms)+'->',*shapes,optimize=opt)[1]\n"
        "    print([(c[0],c[2]) for c in lst])\n"
    )
    outs = set()
    for seed in ("0", "1", "12345"):
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                             env={"PYTHONHASHSEED": seed, "PYTHONPATH": sys.path[0] + ":" + ":".join(sys.path),
                                  "PATH": "/usr/bin:/bin"})
        assert res.returncode == 0, res.stderr
        outs.add(res.stdout)
    assert len(outs) == 1
