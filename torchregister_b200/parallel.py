"""Multi-GPU partitioning of the registration hot path (one process per GPU, torch.distributed).

The reference is single-device; what shards naturally is
  (a) a batch of independent volume pairs  -> split by pair, NO data-path collective;
  (b) one large volume, rigid/affine       -> output z-slabs (2-D: y-slabs) per rank, moving volume
      replicated, ONE all-reduce per epoch of the 41 fp64 moments (328 B), then every rank applies
      the identical update redundantly (bit-identical state on all ranks, no broadcast needed).
The reference-parity flow mode (U-Net parametrised) does not shard spatially: replicas only.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_pairs", "slab_range", "allreduce_moments", "ShardedAffine", "ShardedDirectFlow"]


def shard_pairs(n_pairs: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of pair indices owned by `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, rem = divmod(n_pairs, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def slab_range(n_slices: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of output slices (z in 3-D, y in 2-D) owned by `rank`."""
    return shard_pairs(n_slices, world, rank)


def allreduce_moments(moments: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the [n_pairs, 41] fp64 moment blocks over ranks (NCCL on GPUs, gloo in CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(moments, op=dist.ReduceOp.SUM, group=group)
    return moments


class ShardedAffine:
    """Rigid/affine registration of ONE large pair with the output volume split into slabs over the
    ranks of `group`.  Every rank holds the full moving volume and (for simplicity of addressing) the
    full target; only its own target slab is read."""

    def __init__(self, moving, target, mode, params0, max_epochs, group=None):
        from . import functional as TF
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.prob = TF.AffineProblem(moving, target, mode, params0, max_epochs)
        n_slices = self.prob.D if self.prob.ndim == 3 else self.prob.H
        self.s_begin, self.s_end = slab_range(n_slices, self.world, self.rank)
        self._mom = torch.empty(self.prob.n_pairs, TF.MOMENTS, dtype=torch.float64, device=self.prob.device)

    def run(self, n_epochs, lr, w_mse, w_ncc, optimiser="sgd"):
        for _ in range(n_epochs):
            self.prob.moments(self.s_begin, self.s_end, out=self._mom)
            allreduce_moments(self._mom, self.group)
            self.prob.apply(self._mom, lr, w_mse, w_ncc, optimiser)

    @property
    def final_theta(self):
        return self.prob.final_theta

    @property
    def best_theta(self):
        return self.prob.best_theta

    @property
    def losses(self):
        return self.prob.losses


class ShardedDirectFlow:
    """EXTENSION: direct per-voxel flow registration of ONE large volume split into z-slabs over the ranks
    (BASELINE configs[4]).  Each rank owns the target/flow/optimiser slab [z0, z1); the moving volume is
    replicated.  Per epoch: exchange ONE boundary slice of the flow with each neighbour (the smoothness stencil
    and nothing else crosses slabs), stats pass, all-reduce of the 6 loss moments, update pass."""

    def __init__(self, moving, target, max_epochs, optimiser="sgd", group=None):
        from . import functional as TF
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        D = int(moving.shape[2])
        self.z0, self.z1 = slab_range(D, self.world, self.rank)
        self.prob = TF.DirectFlowProblem(moving, target[:, :, self.z0:self.z1].contiguous(), max_epochs,
                                         z_off=self.z0, optimiser=optimiser)
        nd, H, W = self.prob.ndim, self.prob.H, self.prob.W
        dev = self.prob.device
        self.halo_lo = torch.zeros(nd, H, W, device=dev) if self.rank > 0 else None
        self.halo_hi = torch.zeros(nd, H, W, device=dev) if self.rank < self.world - 1 else None

    def _exchange(self):
        if self.world == 1:
            return
        first, last = self.prob.boundary_slices()
        ops = []
        if self.rank > 0:
            ops += [dist.P2POp(dist.isend, first, self.rank - 1, self.group), dist.P2POp(dist.irecv, self.halo_lo, self.rank - 1, self.group)]
        if self.rank < self.world - 1:
            ops += [dist.P2POp(dist.isend, last, self.rank + 1, self.group), dist.P2POp(dist.irecv, self.halo_hi, self.rank + 1, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def run(self, n_epochs, lr, w_mse, w_ncc, smooth=0.0, betas=(0.9, 0.999), eps=1e-8):
        if self.prob.fused and n_epochs > 0:
            # one kernel + one 6-value all-reduce (+ one halo slice each way) per epoch
            if self.prob.prime(w_ncc):
                allreduce_moments(self.prob.moments, self.group)
            for _ in range(n_epochs):
                if smooth:
                    self._exchange()
                self.prob.step(lr, w_mse, w_ncc, smooth, self.halo_lo, self.halo_hi, betas, eps)
                allreduce_moments(self.prob.moments, self.group)
            self.prob.finish(w_mse, w_ncc, smooth)
            return
        for _ in range(n_epochs):
            if smooth:
                self._exchange()
            m = self.prob.stats(smooth, self.halo_lo, self.halo_hi)
            allreduce_moments(m, self.group)
            self.prob.update(lr, w_mse, w_ncc, smooth, self.halo_lo, self.halo_hi, betas, eps)

    @property
    def flow_slab(self):
        return self.prob.flow

    @property
    def losses(self):
        return self.prob.losses
