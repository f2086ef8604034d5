"""Reverse mode of the plan (contractn_b200/autograd.py) checked on the CPU: the Function is built over torch
stand-ins for the two engine calls it makes, so what is tested is the algebra -- the gradient contractions, the
cotangent of a split-format result, the scale bookkeeping -- against torch's own autograd through the reference's
formulas (einsum -> stabilize, contractn/einsum.py:89-107)."""
import pytest

torch = pytest.importorskip("torch")

from contractn_b200 import autograd, planner  # noqa: E402

CTN_OUT_SPLIT, CTN_OUT_DENSE = 0, 1


def _split(t):
    kappa = t.abs().mean()
    return t / kappa, torch.log(kappa)


def _global(contraction_list, n):
    ins, out = autograd.global_einsum(contraction_list, n)
    return ",".join(ins) + "->" + out


def _fwd(tensors, contraction_list, out_mode):
    t = torch.einsum(_global(contraction_list, len(tensors)), *tensors)
    if out_mode == CTN_OUT_DENSE:
        return t, torch.zeros((), dtype=t.dtype)
    return _split(t)


def _contract_split(einstr, *ops):
    # an off-gauge answer on purpose: any (R_hat, c_r) with R = R_hat exp(c_r) must do
    t = torch.einsum(einstr, *ops)
    return t * 8.0, torch.log(torch.tensor(1.0 / 8.0, dtype=t.dtype))


FN = autograd.make_function(_fwd, _contract_split)

CASES = [
    ("ab,bc,cd->ad", [(3, 4), (4, 5), (5, 6)]),
    ("ab,bc->", [(3, 4), (4, 5)]),                 # scalar result
    ("ab,bc->a", [(3, 4), (4, 5)]),                # c summed inside one operand only
    ("ab,ab,ab->ab", [(3, 4)] * 3),                # copy-node hub (hyperedge)
    ("bi,bj,ij->b", [(6, 3), (6, 4), (3, 4)]),     # batch index
    ("abc->ca", [(2, 3, 4)]),                      # one operand, b summed away
]


@pytest.mark.parametrize("einstr,shapes", CASES)
@pytest.mark.parametrize("dense", [False, True])
def test_gradients_match_torch_autograd(einstr, shapes, dense):
    gen = torch.Generator().manual_seed(7)
    xs = [torch.rand(s, generator=gen, dtype=torch.float64) - 0.3 for s in shapes]
    lst = planner.contract_path(einstr, *shapes)[1]
    w = None

    def loss(out, c):
        nonlocal w
        if w is None:
            w = torch.rand(out.shape, generator=gen, dtype=torch.float64)
        return (out * w).sum() + (0.0 if dense else 3.0 * c)

    a = [x.clone().requires_grad_(True) for x in xs]
    out, c = FN.apply(lst, CTN_OUT_DENSE if dense else CTN_OUT_SPLIT, dense, *a)
    loss(out, c).backward()
    b = [x.clone().requires_grad_(True) for x in xs]
    t = torch.einsum(einstr, *b)
    ref_out, ref_c = (t, torch.zeros(())) if dense else _split(t)
    loss(ref_out, ref_c).backward()
    for ga, gb in zip(a, b):
        assert torch.allclose(ga.grad, gb.grad, rtol=1e-10, atol=1e-12)


def test_only_requested_gradients_are_computed():
    lst = planner.contract_path("ab,bc->ac", (3, 4), (4, 5))[1]
    a = torch.rand(3, 4, dtype=torch.float64, requires_grad=True)
    b = torch.rand(4, 5, dtype=torch.float64)
    out, c = FN.apply(lst, CTN_OUT_SPLIT, False, a, b)
    (out.sum() + c).backward()
    assert a.grad is not None and b.grad is None


def test_global_einsum_is_rebuilt_from_the_pairwise_list():
    shapes = [(3, 4), (4, 5), (5, 6), (6, 2)]
    lst = planner.contract_path("ab,bc,cd,de->ae", *shapes)[1]
    assert autograd.global_einsum(lst, 4) == (["ab", "bc", "cd", "de"], "ae")


def test_label_that_occurs_nowhere_else_becomes_a_ones_vector():
    einstr, spec = autograd.gradient_einsum(["ab", "bc"], "a", 1, [(3, 4), (4, 5)])
    assert einstr == "a,ab,c->bc" and spec[-1] == ("ones", 5)


def test_repeated_label_in_the_differentiated_operand_is_refused():
    with pytest.raises(NotImplementedError):
        autograd.gradient_einsum(["aab", "b"], "", 0, [(3, 3, 4), (4,)])
