"""The reference's derivative recipe: nested ``torch.autograd.grad`` with
``create_graph=True`` (code/pde2d_base.py:46-62, code/ode_base.py:32-42,
code/cavity.py:71-76).  On the fused path these calls do NOT differentiate an
(N, M) graph: they walk the three-node tower of spinn_b200.ops and pick up jets
the CUDA kernel already produced."""
import torch
import torch.autograd as ag


def grad_xy(u, xs, ys):
    return ag.grad(outputs=u, inputs=(xs, ys), grad_outputs=torch.ones_like(u),
                   retain_graph=True, create_graph=True)


def jets_2d(u, xs, ys):
    """u, u_x, u_y, u_xx, u_yy (the mixed terms are computed and dropped, as in the reference)."""
    ux, uy = grad_xy(u, xs, ys)
    uxx, _ = grad_xy(ux, xs, ys)
    _, uyy = grad_xy(uy, xs, ys)
    return u, ux, uy, uxx, uyy


def jets_1d(u, x):
    du = ag.grad(outputs=u, inputs=x, grad_outputs=torch.ones_like(u),
                 retain_graph=True, create_graph=True, allow_unused=True)
    d2u = ag.grad(outputs=du, inputs=x, grad_outputs=torch.ones_like(du[0]),
                  retain_graph=True, create_graph=True, allow_unused=True)
    return u, du[0], d2u[0]
