"""Device-resident interior steps of BASELINE configs 1 and 4 (spinn_b200.fused_step.Ode3InteriorStep,
CavityInteriorStep: forward jets -> residual epilogue kernel -> fused backward, all through the C ABI)
against the reference algorithm (oracle/torch_port.py: dense tensors, nested autograd.grad, backward)
evaluated in fp64 on the same perturbed state."""
import numpy as np
import pytest
import torch

from helpers import mixed_err
from oracle import torch_port as tp

pytestmark = pytest.mark.gpu


def _perturb(nn, seed):
    g = torch.Generator().manual_seed(seed)
    l1, l2 = nn.layer1, nn.layer2
    with torch.no_grad():
        l2.weight.copy_((0.3 * torch.randn(l2.weight.shape, generator=g)).to(l2.weight.device))
        l2.bias.copy_((0.1 * torch.randn(l2.bias.shape, generator=g)).to(l2.weight.device))
        l1.h.mul_((1.0 + 0.3 * torch.rand(l1.h.shape, generator=g)).to(l2.weight.device))


def _d(t):
    return t.detach().cpu().double().requires_grad_(True)


@pytest.mark.parametrize("kind", ["gaussian", "softplus"])
@pytest.mark.parametrize("nodes,samples", [(3, 60), (40, 2001)])
def test_ode3_interior_step_matches_reference_algorithm(kind, nodes, samples):
    from spinn_b200 import common, spinn1d
    from spinn_b200.fused_step import Ode3InteriorStep
    from spinn_b200.problems import ode3
    common.device("cuda")
    try:
        pde = ode3.ODESimple(nodes, samples)
        act = spinn1d.gaussian if kind == "gaussian" else spinn1d.SoftPlus()
        nn = spinn1d.SPINN1D(pde, act).to("cuda")
        _perturb(nn, 5)
        xs = pde.interior()
        step = Ode3InteriorStep(nn, xs)
        flat = step.run(all_reduce=False).clone()
        assert torch.equal(flat, step.run(all_reduce=False))                    # deterministic
        l1, l2 = nn.layer1, nn.layer2
        x, c, h, W, b = _d(xs), _d(l1.center), _d(l1.h), _d(l2.weight), _d(l2.bias)
        xc = torch.cat((c, l1.fixed.detach().cpu().double()))
        loss = tp.ode3_interior_step(kind, x, xc, h, W, b, float(xs.numel()))
        p = step.pack
        assert abs(float(p.view("loss")[0]) - loss) < 2e-5 * abs(loss)
        # (the residual u_xx + f does not depend on the bias: autograd leaves b.grad = None, ours is 0)
        for name, ref in (("d_x", c.grad), ("d_h", h.grad), ("d_W", W.grad), ("d_b", b.grad)):
            got = p.view(name).cpu().numpy()
            want = np.zeros_like(got) if ref is None else ref.numpy().reshape(got.shape)
            assert mixed_err(got, want) < 4e-5, (name, kind)
    finally:
        common.device("cpu")


@pytest.mark.parametrize("kind", ["gaussian", "softplus"])
def test_cavity_interior_step_matches_reference_algorithm(kind):
    from spinn_b200 import common, spinn2d
    from spinn_b200.fused_step import CavityInteriorStep
    from spinn_b200.problems import cavity
    common.device("cuda")
    try:
        pde = cavity.CavityPDE(100, 2500, sample_frac=0.1)
        act = spinn2d.gaussian if kind == "gaussian" else spinn2d.SoftPlus()
        nn = spinn2d.SPINN2D(pde, act).to("cuda")
        _perturb(nn, 6)
        np.random.seed(0)
        xs, ys = pde.interior()                                                  # 250 of 2500 points
        step = CavityInteriorStep(nn, xs, ys)
        flat = step.run(all_reduce=False).clone()
        assert torch.equal(flat, step.run(all_reduce=False))
        l1, l2 = nn.layer1, nn.layer2
        x, y, cx, cy, h, W, b = _d(xs), _d(ys), _d(l1.x), _d(l1.y), _d(l1.h), _d(l2.weight), _d(l2.bias)
        xc = torch.cat((cx, l1.xf.detach().cpu().double()))
        yc = torch.cat((cy, l1.yf.detach().cpu().double()))
        loss = tp.cavity_interior_step(kind, x, y, xc, yc, h, W, b)
        p = step.pack
        assert abs(float(p.view("loss")[0]) - loss) < 2e-5 * abs(loss)
        for name, ref in (("d_x", cx.grad), ("d_y", cy.grad), ("d_h", h.grad), ("d_W", W.grad), ("d_b", b.grad)):
            got = p.view(name).cpu().numpy()
            want = np.zeros_like(got) if ref is None else ref.numpy().reshape(got.shape)
            assert mixed_err(got, want) < 4e-5, (name, kind)
    finally:
        common.device("cpu")
