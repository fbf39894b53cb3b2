"""Random sub-sampling of the interior points (`--sample-frac < 1`), shared by all problem bases.

The reference draws `np.random.choice(rng_interior, sample_size, replace=False)` from the global,
unseeded NumPy stream once per `interior()` call (pde2d_base.py:115-122, ode_base.py:64-68,
ibvp2d_base.py:170-177, fd_spinn1d_base.py:118-125).  The eager loop does exactly that.  In
CUDA-graph mode (spinn_b200/graphed.py) the gather has to read a FIXED device buffer, so the same
draw is made on the host before every replay (`graph_pre_step`) and uploaded into that buffer.
"""
import numpy as np


class SubSampling:
    """Expects `sample_frac`, `rng_interior`, `sample_size` on the instance."""

    _idx_dev = None
    _idx_live = False          # True only while a graphed solve() is replaying against _idx_dev

    def uses_all_samples(self):
        return abs(self.sample_frac - 1.0) < 1e-3

    def draw_indices(self):
        """None = all points; otherwise an index (NumPy array, or the static device buffer)."""
        if self.uses_all_samples():
            return None
        if self._idx_dev is not None and self._idx_live:
            return self._idx_dev
        return np.random.choice(self.rng_interior, size=self.sample_size, replace=False)

    def pick(self, *tensors):
        idx = self.draw_indices()
        return tensors if idx is None else tuple(t[idx] for t in tensors)

    # -- CUDA-graph support; not in the reference ------------------------------------------------
    def _sampling_device(self):
        raise NotImplementedError()

    def enable_static_sampling(self):
        if self.uses_all_samples():
            return
        self._idx_live = True
        if self._idx_dev is not None:          # a later graphed solve(): the captured step reads this buffer
            return
        import torch
        # ring of pinned staging buffers: the host runs ahead of the GPU in graph mode, so a
        # buffer may only be rewritten once the copy that read it has completed
        self._idx_ring = [(torch.empty(self.sample_size, dtype=torch.int64).pin_memory(), torch.cuda.Event())
                          for _ in range(4)]
        self._idx_turn = 0
        self._idx_dev = torch.zeros(self.sample_size, dtype=torch.int64, device=self._sampling_device())

    def disable_static_sampling(self):
        """After a graphed solve(): eager calls of interior() draw fresh indices again (the static
        buffer is kept for the next replay)."""
        self._idx_live = False

    def graph_pre_step(self):
        if self._idx_dev is None or not self._idx_live:
            return
        import torch
        idx = np.random.choice(self.rng_interior, size=self.sample_size, replace=False)
        host, ev = self._idx_ring[self._idx_turn % len(self._idx_ring)]
        if self._idx_turn >= len(self._idx_ring):
            ev.synchronize()
        self._idx_turn += 1
        host.copy_(torch.from_numpy(idx))
        self._idx_dev.copy_(host, non_blocking=True)
        ev.record()
