"""CPU, world_size 2, gloo: the host-side N>1 logic — slab/pair planning and the moment
all-reduce — with the per-slab moments supplied by the CPU oracle (no GPU, no CUDA calls)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle as co
    from torchregister_b200.parallel import allreduce_moments, shard_pairs, slab_range
    from torchregister_b200.synth import make_pair
    # (b) slab sharding: each rank evaluates loss/gradient contributions of its slab only.  The oracle
    # has no slab entry point, so a rank's partial is obtained by zeroing the other slabs' contribution:
    # MSE is additive over voxels, which is exactly the property the sharded path relies on.
    shape = (12, 10, 14)
    mov, tgt = make_pair(shape, "rigid")
    th = np.array([[1.01, 0.02, -0.01, 0.03], [-0.02, 0.99, 0.01, -0.02], [0.01, -0.01, 1.0, 0.01]], np.float64)
    z0, z1 = slab_range(shape[0], world, rank)
    m, t = mov.double().numpy()[0, 0], tgt.double().numpy()[0, 0]
    _, _, warped = co.affine_terms(m, t, th, 1.0, 0.0, want_warped=True)
    sq = ((t - warped) ** 2)[z0:z1].sum()
    part = torch.zeros(1, 41, dtype=torch.float64)
    part[0, 0] = sq
    part[0, 1] = float(z1 - z0)
    allreduce_moments(part)
    full_loss, _, _ = co.affine_terms(m, t, th, 1.0, 0.0)
    ok_slab = abs(part[0, 0].item() / t.size - full_loss) < 1e-12 and part[0, 1].item() == shape[0]
    # (a) batch sharding: disjoint cover, results gathered
    a, b = shard_pairs(5, world, rank)
    mine = torch.zeros(5, dtype=torch.float64)
    mine[a:b] = torch.arange(a, b, dtype=torch.float64) + 1
    dist.all_reduce(mine)
    ok_batch = torch.equal(mine, torch.arange(1, 6, dtype=torch.float64))
    with open(os.path.join(out_dir, "rank%d" % rank), "w") as f:
        f.write("%d %d" % (ok_slab, ok_batch))
    dist.destroy_process_group()


def test_gloo_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / ("rank%d" % r)).read() == "1 1"
