"""The BASELINE 'synthetic scaled Poisson2D' state: what
``poisson2d_sine.py --nodes 3600 --b-nodes 124 --samples 4194304`` would build --
60x60 free + 4x124 fixed nodes (M=4096, h=1/64), a 2048x2048 interior grid
(N=4 194 304), Gaussian kernel, fp32 -- with seeded perturbations of W, h, centres and
bias so that the backward pass is not trivially zero (weights are zero-initialised in
the reference, spinn2d.py:165-167)."""
from __future__ import annotations

import numpy as np
import torch

from . import common
from .problems.poisson2d_sine import Poisson2D
from .spinn2d import SPINN2D, SoftPlus, gaussian


def perturb_(nn, seed=0):
    """Seeded parameter state used by bench/parity tests (CPU generator -> device)."""
    g = torch.Generator().manual_seed(seed)
    l1, l2 = nn.layer1, nn.layer2
    dev = l2.weight.device
    with torch.no_grad():
        l2.weight.copy_((0.3 * torch.randn(l2.weight.shape, generator=g)).to(dev))
        l1.h.mul_((1.0 + 0.3 * torch.rand(l1.h.shape, generator=g)).to(dev))
        l1.x.add_((0.01 * torch.randn(l1.x.shape, generator=g)).to(dev))
        l1.y.add_((0.01 * torch.randn(l1.y.shape, generator=g)).to(dev))
        if l2.bias is not None:
            l2.bias.copy_((0.1 * torch.randn(l2.bias.shape, generator=g)).to(dev))
    return nn


def make_state(nodes=3600, b_nodes=124, samples=4194304, kind="gaussian", device="cuda", seed=0,
               b_samples=None):
    """Returns (pde, nn) on `device`; pde.p_samples are the interior points."""
    prev = common.device()
    common.device(device)
    try:
        pde = Poisson2D(nodes, samples, b_nodes, b_samples)
        act = gaussian if kind == "gaussian" else SoftPlus()
        nn = SPINN2D(pde, act).to(common.device())
        perturb_(nn, seed)
    finally:
        common.device(prev)
    return pde, nn
