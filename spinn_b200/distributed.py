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
