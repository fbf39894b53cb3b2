"""Helper of tests/test_gpu_unchanged_reference.py: execute the reference's problem scripts
BYTE-FOR-BYTE UNCHANGED (from baseline/_ref/code, as `__main__`, with `--gpu`) while
spinn_b200/dropin's `common` / `spinn1d` / `spinn2d` shadow the reference's modules of the same
flat names, so every forward / derivative / backward of those scripts runs on the CUDA kernels.

    python tests/run_unchanged_reference_on_gpu.py <repo root> <out.npz> <script> [script args...]

Writes the loss history, the final error norms and the number of C-ABI forward / backward calls."""
import contextlib
import io
import os
import runpy
import sys

import numpy as np
import torch


def main():
    root, outp, script = sys.argv[1:4]
    extra = sys.argv[4:]
    sys.path.insert(0, root)
    from baseline import reference_loader as rl
    rl.stub_plotting_modules()
    sys.path.insert(0, rl.REF_CODE)                                   # the reference's problem files ...
    sys.path.insert(0, os.path.join(root, "spinn_b200", "dropin"))    # ... our common/spinn1d/spinn2d first
    from spinn_b200 import _cabi, ops
    _cabi.load()
    backend = ops._BACKEND
    assert isinstance(backend, ops.CudaBackend), backend
    calls = {"forward": 0, "backward": 0}
    for name in calls:
        fn = getattr(backend, name)

        def counted(*a, _fn=fn, _name=name, **k):
            calls[_name] += 1
            return _fn(*a, **k)
        setattr(backend, name, counted)

    np.random.seed(0)
    torch.manual_seed(0)
    path = os.path.join(rl.REF_CODE, script + ".py")
    sys.argv = [path] + extra + ["--gpu"]
    with contextlib.redirect_stdout(io.StringIO()):
        g = runpy.run_path(path, run_name="__main__")
    import common
    import spinn2d
    assert common.__file__.startswith(root) and not common.__file__.startswith(rl.REF_CODE), common.__file__
    assert spinn2d.__file__.startswith(os.path.join(root, "spinn_b200")), spinn2d.__file__
    assert g["__file__"].startswith(rl.REF_CODE)
    app = g["app"]
    # the problem / plotter classes are the reference's own, the network and the driver are ours
    assert type(app.pde).__module__ == "__main__" and type(app.nn).__module__.startswith("spinn_b200.")
    assert all(p.is_cuda for p in app.nn.parameters())
    errs = app.plotter.get_error()
    errs = (0.0, 0.0, 0.0) if errs is None else errs
    np.savez(outp, loss=np.asarray(app.solver.loss), errors=np.asarray(errs, dtype=np.float64),
             forward_calls=calls["forward"], backward_calls=calls["backward"])


if __name__ == "__main__":
    main()
