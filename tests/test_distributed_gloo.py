"""world_size-2 gloo test (CPU) of the multi-GPU host logic: point sharding, the packed
[grads|loss] buffer and its single all-reduce.  Per-rank compute is the oracle here (the
product computes it with the CUDA kernels; tests/test_gpu_parity.py checks on the GPU that
shard sums equal the unsharded result)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_case, maxnorm_err


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from oracle import closed_form as cf
    from spinn_b200 import distributed
    r, ws, _ = distributed.init(backend="gloo")
    assert (r, ws) == (rank, world)
    c = case
    N = len(c["x"])
    lo, hi = distributed.shard_range(N, rank, world)
    xc = np.concatenate([c["xf"], c["xfix"]])
    yc = np.concatenate([c["yf"], c["yfix"]])
    Mf, M, K = len(c["xf"]), len(xc), int(c["K"])
    pack = distributed.GradPack(2, Mf, M, K, len(c["h"]), True, torch.device("cpu"))
    g = c["gjets"][:, lo:hi]
    G = cf.grads2d(c["kind"], c["x"][lo:hi], c["y"][lo:hi], xc, yc, c["h"], c["W"], g, n_free=Mf)
    J = cf.jets2d(c["kind"], c["x"][lo:hi], c["y"][lo:hi], xc, yc, c["h"], c["W"], c["b"])
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float32).reshape(-1))
    pack.view("d_x").copy_(t(G["d_xc"]))
    pack.view("d_y").copy_(t(G["d_yc"]))
    pack.view("d_h").copy_(t(G["d_h"]))
    pack.view("d_W").copy_(t(G["d_W"]))
    pack.view("d_b").copy_(t(G["d_b"]))
    pack.view("loss").fill_(float((g * J).sum()))
    pack.all_reduce()
    if rank == 0:
        np.save(out_path, pack.flat.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from spinn_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 4194304, 1001):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_allreduce_equals_unsharded(tmp_path):
    case = load_case("g2d_k3")
    out = str(tmp_path / "flat.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, case, out), nprocs=2, join=True)
    flat = np.load(out)
    from spinn_b200.distributed import GradPack
    c = case
    Mf, M, K = len(c["xf"]), len(c["xf"]) + len(c["xfix"]), int(c["K"])
    pack = GradPack(2, Mf, M, K, len(c["h"]), True, torch.device("cpu"))
    pack.flat.copy_(torch.tensor(flat))
    for name, key in (("d_x", "d_xc"), ("d_y", "d_yc"), ("d_h", "d_h"), ("d_W", "d_W"), ("d_b", "d_b")):
        assert maxnorm_err(pack.view(name).numpy(), np.asarray(c[key + "_f64"]).reshape(-1)) < 1e-6, name
    ref_loss = float((c["gjets"].astype(np.float64) * c["jets_f64"]).sum())
    assert abs(float(pack.view("loss")[0]) - ref_loss) < 1e-5 * abs(ref_loss) + 1e-5


def _train_worker(rank, world, port, out_path, steps):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import contextlib
    import io
    from helpers import OracleBackend
    from spinn_b200 import distributed, ops
    from spinn_b200.problems import poisson2d_sine
    distributed.init(backend="gloo")
    ops.set_backend(OracleBackend())
    np.random.seed(0)
    torch.manual_seed(0)
    app = poisson2d_sine.make_app()
    calls = {"n": 0}
    real_all_reduce = dist.all_reduce

    def counting_all_reduce(*a, **k):
        calls["n"] += 1
        return real_all_reduce(*a, **k)
    dist.all_reduce = counting_all_reduce
    with contextlib.redirect_stdout(io.StringIO()):
        app.run(args=["--nodes", "100", "--samples", "400", "--n-train", str(steps), "--lr", "1e-3",
                      "--tol", "1e-30"])
    dist.all_reduce = real_all_reduce
    from spinn_b200.dist_optimizer import DistributedOptimizer
    assert isinstance(app.solver, DistributedOptimizer)
    assert app.pde.p_samples[0].numel() == 400 // world
    # boundary ring (76 points) sharded like the interior; ONE collective per closure, gradients are
    # views of the reduced buffer (no pack / unpack)
    assert app.solver.boundary_sharded and app.pde.b_samples[0].numel() == 76 // world
    assert calls["n"] == steps, (calls, steps)
    flat = app.solver._flat
    assert all(p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() for p in app.nn.parameters())
    if rank == 0:
        np.save(out_path, np.asarray(app.solver.loss))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_training_loop_reproduces_single_process_losses(tmp_path):
    """App.run under WORLD_SIZE=2 (gloo, oracle-backed op): sharded interior + one gradient
    all-reduce per step gives the single-process loss history (reference-recorded)."""
    steps = 25
    out = str(tmp_path / "loss.npy")
    mp.spawn(_train_worker, args=(2, _free_port(), out, steps), nprocs=2, join=True)
    loss = np.load(out)
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "traj_poisson2d_sine.npz"))["loss"]
    rel = np.abs(loss - ref[:steps]) / np.abs(ref[:steps])
    assert rel.max() < 5e-5, rel.max()


def _fd_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import contextlib
    import io
    from helpers import OracleBackend
    from spinn_b200 import distributed, ops
    from spinn_b200.problems import heat1d_fd_spinn
    distributed.init(backend="gloo")
    ops.set_backend(OracleBackend())
    np.random.seed(0)
    torch.manual_seed(0)
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "traj_heat1d_fd.npz"))
    app = heat1d_fd_spinn.make_app()
    with contextlib.redirect_stdout(io.StringIO()):
        app.run(args=[str(a) for a in d["args"]])
    # the carried field follows the local slice of the sample points
    assert app.pde.xs.numel() == 100 // world and app.pde.u0.shape == app.pde.xs.shape
    if rank == 0:
        np.save(out_path, np.asarray(app.solver.loss))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_fd_spinn_time_stepping_reproduces_reference(tmp_path):
    """FD-SPINN outer loop under WORLD_SIZE=2: the previous-level field u0 is sharded with the
    sample points, every level's solve() is data-parallel, tolerance stops agree on all ranks."""
    out = str(tmp_path / "loss.npy")
    mp.spawn(_fd_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    loss = np.load(out)
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "traj_heat1d_fd.npz"))["loss"]
    assert len(loss) == len(ref)
    rel = np.abs(loss - ref) / np.abs(ref)
    assert rel.max() < 2e-4, rel.max()
