"""Reverse mode through the device plan against torch's autograd through the reference's own formulas
(einsum -> stabilize, contractn/einsum.py:89-107, 326-393), on the GPU, through the public `contract` and
`TN.contract` calls.  float64: 1e-10 relative; float32: 2e-4 (gradients are contractions of the same size)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from contractn_b200 import engine, networks  # noqa: E402

pytestmark = pytest.mark.gpu

CASES = [
    ("ab,bc,cd->ad", [(30, 40), (40, 50), (50, 60)]),
    ("ab,bc->", [(30, 40), (40, 50)]),
    ("ab,bc->a", [(30, 40), (40, 50)]),
    ("ab,ab,ab->ab", [(64, 48)] * 3),
    ("bi,bj,ij->b", [(600, 30), (600, 40), (30, 40)]),
    ("br,prs,bp->bs", [(256, 64), (2, 64, 96), (256, 2)]),   # one cfg-3 site (tensor-core step for float32)
    ("abc->ca", [(20, 30, 40)]),
]


def _ref_split(t):
    kappa = t.abs().mean()
    return t / kappa, torch.log(kappa)


@pytest.mark.parametrize("einstr,shapes", CASES)
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("split", [True, False])
def test_contract_gradients(einstr, shapes, dtype, split):
    gen = torch.Generator().manual_seed(11)
    xs = [(torch.rand(s, generator=gen, dtype=torch.float64) + 0.1).to(dtype).cuda() for s in shapes]
    a = [x.clone().requires_grad_(True) for x in xs]
    b = [x.double().clone().requires_grad_(True) for x in xs]
    ref = torch.einsum(einstr, *b)
    w = torch.rand(ref.shape, generator=gen, dtype=torch.float64).cuda()
    if split:
        out, c = engine.contract(einstr, *a, split_format=True)
        loss = (out.double() * w).sum() + 3.0 * c.double()
        r_out, r_c = _ref_split(ref)
        ref_loss = (r_out * w).sum() + 3.0 * r_c
    else:
        out = engine.contract(einstr, *a)
        loss = (out.double() * w).sum()
        ref_loss = (ref * w).sum()
    loss.backward()
    ref_loss.backward()
    rtol = 1e-10 if dtype == torch.float64 else 2e-4
    assert abs(loss.item() - ref_loss.item()) <= rtol * abs(ref_loss.item()) * 10
    for ga, gb in zip(a, b):
        assert ga.grad is not None and ga.grad.dtype == dtype
        scale = gb.grad.abs().max()
        assert float((ga.grad.double() - gb.grad).abs().max()) <= rtol * float(scale)


def test_tn_contract_backpropagates():
    """A batched MPS (cfg 3's network at toy size) through TN.contract: gradients w.r.t. every core."""
    tn, path = networks.mps_tn(sites=6, bond=8, phys=2, batch=16, dtype=np.float64, seed=5)
    params = [torch.as_tensor(np.asarray(p)).cuda().requires_grad_(True) for p in tn.params]
    out, c = engine.contract(tn.einsum_str, *params, optimize=path, split_format=True)
    (out.sum() + c).backward()
    ref_params = [p.detach().clone().requires_grad_(True) for p in params]
    r_out, r_c = _ref_split(torch.einsum(tn.einsum_str, *ref_params))
    (r_out.sum() + r_c).backward()
    for p, q in zip(params, ref_params):
        assert torch.allclose(p.grad, q.grad, rtol=1e-9, atol=1e-12 * float(q.grad.abs().max()))


def test_no_grad_and_detached_calls_take_the_plain_path():
    a = torch.rand(8, 9, dtype=torch.float64, device="cuda", requires_grad=True)
    b = torch.rand(9, 7, dtype=torch.float64, device="cuda")
    with torch.no_grad():
        out, c = engine.contract("ab,bc->ac", a, b, split_format=True)
    assert not out.requires_grad and not c.requires_grad
    out, c = engine.contract("ab,bc->ac", a, b, split_format=True)
    assert out.requires_grad and c.requires_grad
