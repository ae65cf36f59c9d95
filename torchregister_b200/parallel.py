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


class PeerMailbox:
    """Peer-mapped mailbox for the fused all-reduce of the sharded affine epoch (include/trb.h: trb_affine_optim_peer).
    Every rank allocates 2*8*48+8 float64 in symmetric memory (torch.distributed._symmetric_memory: the buffer of every
    rank is mapped into every process, peer access over NVLink); `ptrs[r]` is rank r's buffer in this process."""

    DOUBLES = 2 * 8 * 48 + 8             # slots + the poison word (csrc/peer.cuh)

    def __init__(self, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(grp), dist.get_rank(grp)
        if self.world > 8:
            raise ValueError("the peer mailbox holds up to 8 ranks (one NVSwitch box)")
        self.buf = symm_mem.empty(self.DOUBLES, dtype=torch.float64, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, grp)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        torch.cuda.synchronize(device)
        dist.barrier(grp)                       # every mailbox is zeroed and mapped before the first push
        self.seq = 1

    def take(self, n):
        s = self.seq
        self.seq += int(n)
        return s

    def check(self):
        """Raise if an in-kernel exchange gave up waiting for a peer (csrc/peer.cuh: bounded spin, then the moments
        are poisoned with NaN and this word is set).  Synchronises the device; call it after the epochs of interest."""
        poison = self.buf[2 * 8 * 48:2 * 8 * 48 + 1].view(torch.int64)
        if int(poison.item()) != 0:
            raise RuntimeError(
                "peer exchange timed out on rank %d: a rank of the group did not reach the same epoch within the spin "
                "budget (results from that epoch on are NaN). Re-create the sharded problem after fixing the skew." % self.rank)

    def reset(self):
        """Clear the poison word and the slots (collective: every rank, then a barrier)."""
        torch.cuda.synchronize(self.buf.device)
        self.buf.zero_()
        torch.cuda.synchronize(self.buf.device)
        dist.barrier()
        self.seq = 1


class ShardedAffine:
    """Rigid/affine registration of ONE large pair with the output volume split into slabs over the
    ranks of `group`.  Every rank holds the full moving volume and (for simplicity of addressing) the
    full target; only its own target slab is read."""

    def __init__(self, moving, target, mode, params0, max_epochs, group=None, peer=None):
        from . import functional as TF
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.prob = TF.AffineProblem(moving, target, mode, params0, max_epochs)
        n_slices = self.prob.D if self.prob.ndim == 3 else self.prob.H
        self.s_begin, self.s_end = slab_range(n_slices, self.world, self.rank)
        self._mom = torch.empty(self.prob.n_pairs, TF.MOMENTS, dtype=torch.float64, device=self.prob.device)
        # fused form (one kernel per epoch, moments all-reduced inside it through peer memory) when the box allows
        # it; otherwise moments kernel -> NCCL all-reduce -> apply kernel
        self.mailbox, self.peer_error = None, None
        if peer is not False and self.world > 1 and self.prob.ndim == 3 and self.prob.n_pairs == 1:
            ok = 1
            try:
                self.mailbox = PeerMailbox(self.prob.device, group)
                self.prob.run_peer(0, self.s_begin, self.s_end, self.mailbox.ptrs, self.rank, self.world, 1, 0.0, 0.0, 1.0)
            except Exception as e:                       # no symmetric memory on this box / shape the TMA kernel rejects
                ok, self.peer_error = 0, repr(e)
            flag = torch.tensor([ok], device=self.prob.device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)      # all ranks take the same path
            if int(flag.item()) == 0:
                self.mailbox = None
                if peer is True:
                    raise RuntimeError("fused sharded epoch unavailable: %s" % (self.peer_error or "another rank declined"))

    def run(self, n_epochs, lr, w_mse, w_ncc, optimiser="sgd", check=False, align=True):
        if self.mailbox is not None:
            if align:
                # ranks may arrive here seconds apart (data loading, lazy module loads): line them up first, the
                # in-kernel exchange only spins for a bounded time (align=False: the caller has just done so)
                torch.cuda.current_stream(self.prob.device).synchronize()
                dist.barrier(self.group)
            self.prob.run_peer(n_epochs, self.s_begin, self.s_end, self.mailbox.ptrs, self.rank, self.world,
                               self.mailbox.take(n_epochs), lr, w_mse, w_ncc, optimiser)
            if check:
                self.mailbox.check()
            return
        for _ in range(n_epochs):
            self.prob.moments(self.s_begin, self.s_end, out=self._mom)
            allreduce_moments(self._mom, self.group)
            self.prob.apply(self._mom, lr, w_mse, w_ncc, optimiser)

    def check(self):
        """Raise if the fused exchange timed out on a peer (see PeerMailbox.check)."""
        if self.mailbox is not None:
            self.mailbox.check()

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

    def __init__(self, moving, target, max_epochs, optimiser="sgd", group=None, peer=None):
        from . import functional as TF
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        D = int(moving.shape[2])
        self.z0, self.z1 = slab_range(D, self.world, self.rank)
        dev = moving.device
        nd = moving.dim() - 2
        H, W = int(moving.shape[-2]), int(moving.shape[-1])
        # Fused form (default when the box offers symmetric memory): the ping-pong flow buffers of every rank are mapped
        # into every process, so a rank reads its neighbours' boundary slices in place and the epoch kernel all-reduces
        # its 6 sums through the peer mailboxes — ONE kernel per epoch, no NCCL call.  Otherwise: halo slices by
        # batch_isend_irecv, fused kernel, NCCL all-reduce.
        self.mailbox, self.peer_error, self._peer_flow = None, None, None
        buffers = None
        if peer is not False and self.world > 1 and nd == 3:
            ok = 1
            try:
                import torch.distributed._symmetric_memory as symm_mem
                grp = group if group is not None else dist.group.WORLD
                counts = [3 * (slab_range(D, self.world, r)[1] - slab_range(D, self.world, r)[0]) * H * W for r in range(self.world)]
                n = max(counts)                                        # symmetric allocations have one size on all ranks
                bufs = [symm_mem.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
                hdls = [symm_mem.rendezvous(b, grp) for b in bufs]
                self._peer_flow = [[int(p) for p in h.buffer_ptrs] for h in hdls]
                self._peer_keep = (bufs, hdls)
                mine = 3 * (self.z1 - self.z0) * H * W
                buffers = [b[:mine] for b in bufs]
                self.mailbox = PeerMailbox(dev, group)
            except Exception as e:
                ok, self.peer_error = 0, repr(e)
            flag = torch.tensor([ok], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)      # all ranks take the same path
            if int(flag.item()) == 0:
                self.mailbox, buffers = None, None
                if peer is True:
                    raise RuntimeError("fused sharded flow unavailable: %s" % (self.peer_error or "another rank declined"))
        self.prob = TF.DirectFlowProblem(moving, target[:, :, self.z0:self.z1].contiguous(), max_epochs,
                                         z_off=self.z0, optimiser=optimiser, flow_buffers=buffers)
        self._hw = H * W
        self._ds = [slab_range(D, self.world, r)[1] - slab_range(D, self.world, r)[0] for r in range(self.world)]
        self._phase = 0                                                 # which of the two symmetric buffers is `flow`
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

    def _peer_halos(self):
        """Raw pointers to the neighbours' boundary slices inside their CURRENT flow buffer, and their channel strides."""
        cur = self._peer_flow[self._phase]
        lo = hi = None
        lo_cs = hi_cs = 0
        if self.rank > 0:
            ds = self._ds[self.rank - 1]
            lo, lo_cs = cur[self.rank - 1] + 4 * (ds - 1) * self._hw, ds * self._hw      # its last slice, channel 0
        if self.rank < self.world - 1:
            hi, hi_cs = cur[self.rank + 1], self._ds[self.rank + 1] * self._hw            # its first slice
        return lo, lo_cs, hi, hi_cs

    def run(self, n_epochs, lr, w_mse, w_ncc, smooth=0.0, betas=(0.9, 0.999), eps=1e-8):
        if self.mailbox is not None and n_epochs > 0:
            if self.prob.prime(w_ncc):
                allreduce_moments(self.prob.moments, self.group)
            torch.cuda.current_stream(self.prob.device).synchronize()
            dist.barrier(self.group)            # every rank's current flow is in place before a neighbour reads it
            for _ in range(n_epochs):
                lo, lo_cs, hi, hi_cs = self._peer_halos()
                self.prob.step_peer(lr, w_mse, w_ncc, smooth, lo, lo_cs, hi, hi_cs, self.mailbox.ptrs, self.rank, self.world,
                                    self.mailbox.take(1), betas, eps)
                self._phase ^= 1
            self.prob.finish(w_mse, w_ncc, smooth)
            return
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

    def check(self):
        """Raise if the fused exchange timed out on a peer (see PeerMailbox.check)."""
        if self.mailbox is not None:
            self.mailbox.check()

    @property
    def flow_slab(self):
        return self.prob.flow

    @property
    def losses(self):
        return self.prob.losses
