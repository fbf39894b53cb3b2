"""Point sharding across GPUs + the one collective of a training step.

The reference is single-process (no collective call sites);
this is new.  Loss and every parameter gradient of the SPINN residual are plain
sums over collocation points, so rank r owns a contiguous slice of the points,
node parameters (~80 KB) are replicated, each rank computes its local
[d_x | d_y | d_h | d_W | d_b | loss] vector (scaled by 1/N_global), and ONE
``all_reduce(SUM)`` (NCCL over NVLink/NVSwitch) makes gradients and loss identical
on all ranks; the optimizer step is then replicated.  One process per GPU,
launched by torchrun; ``gloo`` works too (CPU tests of this logic).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend=None):
    """Initialise torch.distributed from the environment when WORLD_SIZE > 1."""
    rank, ws, local = world()
    if ws > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, ws, local


def shard_range(n, rank, world_size):
    """Contiguous, balanced slice [lo, hi) of n points for `rank`; sizes differ by <= 1
    and are kept even where possible (the backward kernel packs points in pairs)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


class GradPack:
    """Flat fp32 buffer [d_x(Mf) | d_y(Mf) | d_h(Mh) | d_W(K*M) | d_b(K) | loss(1)] whose
    slices are handed to the C ABI as output pointers, so the backward's second stage
    writes straight into the message that is all-reduced."""

    def __init__(self, dim, m_free, m_total, k, h_numel, has_bias, device):
        self.dim = dim
        sizes = [("d_x", m_free)]
        if dim == 2:
            sizes.append(("d_y", m_free))
        sizes += [("d_h", h_numel), ("d_W", k * m_total)]
        if has_bias:
            sizes.append(("d_b", k))
        sizes.append(("loss", 1))
        self.slices = {}
        off = 0
        for name, n in sizes:
            # keep every segment 16-byte aligned for vectorised writers
            self.slices[name] = (off, n)
            off += (n + 3) // 4 * 4
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.k, self.m_total = k, m_total

    def view(self, name):
        off, n = self.slices[name]
        return self.flat[off:off + n]

    def has(self, name):
        return name in self.slices

    def all_reduce(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        return self.flat

    def nbytes(self):
        return self.flat.numel() * 4


# --------------------------------------------------------------------------------------
# Data-parallel training through the plug-in surface (any problem class, any config)
# --------------------------------------------------------------------------------------
def shard_interior(pde, rank=None, world_size=None):
    """Make `pde.interior()` return this rank's contiguous slice of the interior points.

    Works for problem classes that keep their interior samples in `p_samples`
    (RegularPDE / IBVP2D / CavityPDE), `interior_samples` (irregular-domain Poisson2D) or
    `xs` (BasicODE).  Every rank keeps ALL boundary points; shards must be equal-sized (checked)
    so that the mean over ranks of local means is the global mean.  DistributedOptimizer combines
    the two loss terms accordingly.  Random sub-sampling (`sample_frac<1`) draws from the local
    slice with the local RNG stream."""
    if rank is None:
        rank, world_size, _ = world()
    if world_size == 1:
        return pde
    if getattr(pde, "_shard", (None, None))[:2] == (rank, world_size):
        return pde                                     # already holds this rank's slice
    for attr in ("p_samples", "interior_samples", "xs"):
        if hasattr(pde, attr):
            break
    else:
        raise TypeError("do not know where this problem class keeps its interior samples")
    pts = getattr(pde, attr)
    single = torch.is_tensor(pts)
    tensors = (pts,) if single else tuple(pts)
    n = tensors[0].numel()
    if n % world_size != 0:
        raise ValueError(f"{n} interior points do not split evenly over {world_size} ranks")
    lo, hi = shard_range(n, rank, world_size)
    local = tuple(t.detach()[lo:hi].clone().requires_grad_(t.requires_grad) for t in tensors)
    setattr(pde, attr, local[0] if single else local)
    if hasattr(pde, "rng_interior"):
        import numpy as np
        pde.n_interior_local = hi - lo
        pde.rng_interior = np.arange(hi - lo)
        pde.sample_size = int(pde.sample_frac * (hi - lo))
    pde._shard = (rank, world_size, lo, hi)
    if hasattr(pde, "on_shard"):            # per-point state that follows the slice (FD-SPINN's u0)
        pde.on_shard()
    return pde


def shard_boundary(pde, rank=None, world_size=None, min_per_rank=32):
    """Shard the boundary points like the interior when the problem class says its
    boundary penalty is a plain SUM over boundary points (`boundary_reduction = "sum"`, e.g. the
    Dirichlet penalties of the Poisson configs) and keeps them in `b_samples`.  Returns True if the
    boundary is now sharded (local sums then add up across ranks), False if every rank keeps all of
    it.  Cached boundary targets are dropped so that they are rebuilt for the slice."""
    if rank is None:
        rank, world_size, _ = world()
    if world_size == 1 or getattr(pde, "boundary_reduction", None) != "sum":
        return False
    if getattr(pde, "_bshard", None) == (rank, world_size):
        return True
    pts = getattr(pde, "b_samples", None)
    if pts is None or torch.is_tensor(pts) or len(pts) == 0:
        return False
    n = pts[0].numel()
    if n < min_per_rank * world_size:
        return False
    lo, hi = shard_range(n, rank, world_size)
    pde.b_samples = tuple(t.detach()[lo:hi].clone().requires_grad_(t.requires_grad) for t in pts)
    for cached in ("_ub", "_ub_dev"):
        if hasattr(pde, cached):
            setattr(pde, cached, None)
    pde._bshard = (rank, world_size)
    return True


def all_reduce_gradients(params, average=True, buf=None):
    """ONE collective for all parameter gradients: pack -> all_reduce -> unpack."""
    params = [p for p in params if p.grad is not None]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or not params:
        return buf
    n = sum(p.grad.numel() for p in params)
    if buf is None or buf.numel() != n or buf.device != params[0].grad.device:
        buf = torch.empty(n, dtype=params[0].grad.dtype, device=params[0].grad.device)
    off = 0
    for p in params:
        k = p.grad.numel()
        buf[off:off + k].copy_(p.grad.reshape(-1))
        off += k
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    if average:
        buf.div_(dist.get_world_size())
    off = 0
    for p in params:
        k = p.grad.numel()
        p.grad.copy_(buf[off:off + k].reshape(p.grad.shape))
        off += k
    return buf
