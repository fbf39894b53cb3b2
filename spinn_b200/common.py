"""Driver / plug-in surface, mirroring the reference's ``common`` module.

Same public names, argument meaning and behaviour as /root/reference/code/common.py
(``device`` :14-21, ``tensor`` :24-28, ``PDE`` :31-92, ``Plotter`` :94-151,
``Optimizer`` :154-314, ``App`` :317-379) so that problem scripts written against
the reference drive the B200 path unchanged.  Written from the behaviour, not
from the text: flags are declared in tables and the training loop is split into
small methods.  One flag is new: ``--graph`` replays the training step as a CUDA
graph (spinn_b200/graphed.py); without it the loop is the reference's.
"""
from __future__ import annotations

import os
import time
from argparse import ArgumentDefaultsHelpFormatter, ArgumentParser

import numpy as np
import torch

_state = {"device": torch.device("cpu")}


def device(dev=None):
    """Get (no argument) or set the module-global device (reference common.py:14-21)."""
    if dev is None:
        return _state["device"]
    _state["device"] = torch.device(dev)


def tensor(x, **kw):
    """fp32 tensor on the current global device (reference common.py:24-28)."""
    return torch.tensor(x, dtype=torch.float32, device=device(), **kw)


def add_flags(parser, table, kw):
    """Declare argparse flags from rows ``(flags, dest, default, type, help)``.
    ``kw`` overrides defaults by ``dest`` exactly like ``kw.get(dest, default)`` in
    the reference's setup_argparse methods; ``type`` None means a store_true flag."""
    for flags, dest, default, typ, text in table:
        if typ is None:
            parser.add_argument(*flags, dest=dest, action="store_true",
                                default=kw.get(dest, default), help=text)
        elif isinstance(typ, (list, tuple)):
            parser.add_argument(*flags, dest=dest, default=kw.get(dest, default),
                                choices=list(typ), help=text)
        else:
            parser.add_argument(*flags, dest=dest, default=kw.get(dest, default),
                                type=typ, help=text)


class PDE:
    """Problem protocol (reference common.py:31-92).  Subclasses provide geometry
    (nodes / fixed_nodes / interior / boundary / plot_points), the physics
    (pde / exact / interior_loss / boundary_loss) and the CLI hooks."""

    @classmethod
    def from_args(cls, args):
        return None

    @classmethod
    def setup_argparse(cls, parser, **kw):
        return None

    def nodes(self):
        return None

    def fixed_nodes(self):
        return None

    def interior(self):
        return None

    def boundary(self):
        return None

    def plot_points(self):
        return None

    def n_vars(self):
        return 1

    def pde(self, *args):
        raise NotImplementedError()

    def has_exact(self):
        return True

    def exact(self, *args):
        return None

    def interior_loss(self, nn):
        raise NotImplementedError()

    def boundary_loss(self, nn):
        raise NotImplementedError()

    def loss(self, nn):
        # interior first, boundary second: the order matters for unseeded sampling
        li = self.interior_loss(nn)
        lb = self.boundary_loss(nn)
        return li + lb

    def _compute_derivatives(self, u, x):
        raise NotImplementedError()


NO_LIVE_PLOT = ("live plotting (--plot) is out of scope of spinn_b200 (SURVEY.md section 2, row 11: "
                "visualisation): run without --plot, or plug in your own Plotter subclass; error norms "
                "(get_error) and result files (save) are provided")


class Plotter:
    """Plot/err protocol (reference common.py:94-151); plotting itself is optional."""

    FLAGS = ((("--no-show-exact",), "no_show_exact", False, None,
              "Do not show exact solution even if available."),)

    @classmethod
    def from_args(cls, pde, nn, args):
        return cls(pde, nn, args.no_show_exact)

    @classmethod
    def setup_argparse(cls, parser, **kw):
        add_flags(parser, cls.FLAGS, kw)

    def __init__(self, pde, nn, no_show_exact=False):
        self.pde = pde
        self.nn = nn
        self.plt1 = None
        self.plt2 = None
        self.show_exact = not no_show_exact

    def get_plot_data(self):
        return None

    def get_error(self, **kw):
        return None

    def plot_solution(self):
        return None

    def plot(self):
        return None

    def plot_weights(self):
        return None

    def save(self, dirname):
        return None

    def show(self):
        return None


class Optimizer:
    """Training loop (reference common.py:154-314): ``opt.step(closure)`` for
    ``n_train`` iterations, error report every ``n_skip``, tolerance stop that
    takes effect one iteration after it is detected (common.py:267-268,288-293),
    ``solver.npz`` on save (common.py:300-314)."""

    FLAGS = (
        (("--n-train", "-t"), "n_train", 2500, int, "Number of training iterations."),
        (("--n-skip",), "n_skip", 100, int, "Number of iterations after which we print/update plots."),
        (("--tol", "-e"), "tol", 1e-6, float, "Tolerance for loss computation."),
        (("--plot",), "plot", False, None, "Show a live plot of the results."),
        (("--lr",), "lr", 1e-2, float, "Learning rate."),
        (("--optimizer",), "optimizer", "Adam", ("Adam", "LBFGS"), "Optimizer to use."),
        (("-d", "--directory"), "directory", None, str, "Output directory (output files are dumped here)."),
        # not in the reference: capture one training step in a CUDA graph (spinn_b200/graphed.py)
        (("--graph",), "graph", False, None, "Replay the training step as a CUDA graph (Adam, --gpu)."),
    )

    @classmethod
    def from_args(cls, pde, nn, plotter, args):
        opt_class = {"Adam": torch.optim.Adam, "LBFGS": torch.optim.LBFGS}[args.optimizer]
        if getattr(args, "graph", False) and cls is Optimizer:
            from .graphed import GraphedOptimizer                 # also data-parallel under torchrun
            cls = GraphedOptimizer
        elif cls is Optimizer and int(os.environ.get("WORLD_SIZE", "1")) > 1:
            from .dist_optimizer import DistributedOptimizer      # one process per GPU (torchrun)
            cls = DistributedOptimizer
        return cls(pde, nn, plotter, n_train=args.n_train, n_skip=args.n_skip, tol=args.tol,
                   lr=args.lr, plot=args.plot, out_dir=args.directory, opt_class=opt_class)

    @classmethod
    def setup_argparse(cls, parser, **kw):
        # --plot is a pure CLI switch in the reference (kw does not override it)
        add_flags(parser, cls.FLAGS, {k: v for k, v in kw.items() if k not in ("plot", "graph")})

    def __init__(self, pde, nn, plotter, n_train, n_skip=100, tol=1e-6, lr=1e-2, plot=True,
                 out_dir=None, opt_class=torch.optim.Adam):
        self.pde, self.nn, self.plotter = pde, nn, plotter
        self.opt_class = opt_class
        self.errors_L1, self.errors_L2, self.errors_Linf, self.loss = [], [], [], []
        self.time_taken = 0.0
        self.n_train, self.n_skip, self.tol, self.lr = n_train, n_skip, tol, lr
        self.plot, self.out_dir = plot, out_dir
        self.opt = None
        self.loss_for_tolerance = False

    def make_optimizer(self):
        """`opt_class(nn.parameters(), lr=lr)` as the reference (common.py:255); Adam on CUDA parameters
        is asked for its single multi-tensor kernel (`fused=True`: same update rule, one launch instead
        of ~10 per step -- the small configs are launch-bound)."""
        params = list(self.nn.parameters())
        if self.opt_class is torch.optim.Adam and params and all(
                p.is_cuda and p.dim() > 0 and p.dtype == torch.float32 for p in params):
            return torch.optim.Adam(params, lr=self.lr, fused=True)
        return self.opt_class(params, lr=self.lr)

    def use_loss_for_tolerance(self):
        self.loss_for_tolerance = True

    def closure(self):
        self.opt.zero_grad()
        loss = self.pde.loss(self.nn)
        # retain_graph: problems such as the cavity re-use a cat() of leaf tensors
        # built once in __init__ (reference common.py:246, cavity.py:62-65)
        loss.backward(retain_graph=True)
        self.loss.append(loss.item())
        return loss

    def _report(self, i, loss):
        e1 = e2 = einf = 0.0
        errs = self.plotter.plot() if self.plot else self.plotter.get_error()
        if errs is not None:
            e1, e2, einf = errs
        self.errors_L1.append(e1)
        self.errors_L2.append(e2)
        self.errors_Linf.append(einf)
        lval = loss.item()
        extra = f", Linf error={einf:.3e}" if self.pde.has_exact() else ""
        print(f"Iteration ({i}/{self.n_train}): Loss={lval:.3e}" + extra)
        return lval if (abs(einf) < 1e-8 or self.loss_for_tolerance) else einf

    def solve(self):
        if self.opt is None:
            self.opt = self.make_optimizer()
        if self.plot:
            self.plotter.plot()
        t0 = time.perf_counter()
        err, stop = 1.0, False
        for i in range(1, self.n_train + 1):
            loss = self.opt.step(self.closure)
            if err < self.tol:
                stop = True
            if i % self.n_skip == 0 or i == self.n_train or stop:
                err = self._report(i, loss)
            if stop:
                break
        self.time_taken = time.perf_counter() - t0
        print(f"Done. Took {self.time_taken:.3f} seconds.")
        if self.plot:
            self.plotter.show()

    def save(self):
        if self.out_dir is None:
            print("No output directory set.  Skipping.")
            return
        os.makedirs(self.out_dir, exist_ok=True)
        print("Saving output to", self.out_dir)
        np.savez(os.path.join(self.out_dir, "solver.npz"), loss=self.loss,
                 error_L1=self.errors_L1, error_L2=self.errors_L2,
                 error_Linf=self.errors_Linf, time_taken=self.time_taken)
        self.plotter.save(self.out_dir)


class App:
    """Composition root (reference common.py:317-379): ``App(pde_cls, nn_cls,
    plotter_cls, optimizer).run(args=None, **kw)``."""

    def __init__(self, pde_cls, nn_cls, plotter_cls, optimizer=Optimizer):
        self.pde_cls, self.nn_cls = pde_cls, nn_cls
        self.plotter_cls, self.optimizer = plotter_cls, optimizer

    def _setup_activation_options(self, p, **kw):
        return None

    def _get_activation(self, args):
        return None

    def setup_argparse(self, **kw):
        p = ArgumentParser(description="Configurable options.",
                           formatter_class=ArgumentDefaultsHelpFormatter)
        p.add_argument("--gpu", dest="gpu", action="store_true", default=kw.get("gpu", False),
                       help="Run code on the GPU.")
        for c in (self.pde_cls, self.nn_cls, self.plotter_cls, self.optimizer):
            c.setup_argparse(p, **kw)
        # like the reference (common.py:352) the activation options do NOT see **kw
        self._setup_activation_options(p)
        return p

    def run(self, args=None, **kw):
        args = self.setup_argparse(**kw).parse_args(args)
        if args.gpu and int(os.environ.get("WORLD_SIZE", "1")) > 1:
            device(f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}")   # torchrun: one GPU per process
        else:
            device("cuda" if args.gpu else "cpu")
        activation = self._get_activation(args)
        self.pde = self.pde_cls.from_args(args)
        self.nn = self.nn_cls.from_args(self.pde, activation, args).to(device())
        self.plotter = self.plotter_cls.from_args(self.pde, self.nn, args)
        self.solver = self.optimizer.from_args(self.pde, self.nn, self.plotter, args)
        self.solver.solve()
        if args.directory is not None:
            self.solver.save()
