import sys, io, contextlib, numpy as np, torch
sys.path.insert(0,'/root/repo')
from spinn_b200 import common
from spinn_b200.problems import poisson2d_irreg_dom as irr, accuracy
res=[]
for k in range(-6,7):
    np.random.seed(0); torch.manual_seed(0)
    app = irr.make_app()
    p = app.setup_argparse().parse_args(["--lr","1e-3","--gpu","--graph","--n-skip","100000"])
    common.device("cuda")
    pde = app.pde_cls.from_args(p); nn = app.nn_cls.from_args(pde, app._get_activation(p), p).to("cuda")
    with torch.no_grad(): nn.layer1.h.mul_(1.0 + k*1e-7)
    plotter = app.plotter_cls.from_args(pde, nn, p); solver = app.optimizer.from_args(pde, nn, plotter, p)
    with contextlib.redirect_stdout(io.StringIO()): solver.solve()
    e = accuracy.fem_errors(nn)
    res.append((k, solver.loss[-1], e[2])); print(k, solver.loss[-1], e, flush=True)
