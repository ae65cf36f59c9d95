"""Multi-GPU checks (run under torchrun on N >= 2 GPUs of one box; wrapped by test_multi_gpu.py):
the sharded forms must reproduce the single-GPU result computed on the same rank's GPU."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_checks(rank, world, dev):
    """The checks proper; needs an initialised NCCL process group.  Returns the summary string (also used by
    bench.py --gpus 2, which records it in its JSON line)."""
    import torchregister_b200.functional as TF
    from torchregister_b200.parallel import ShardedAffine, ShardedDirectFlow, shard_pairs
    from torchregister_b200.synth import make_pair, smooth_flow

    # (b) one volume, rigid, z-slabs + all-reduce of the 41 moments
    mov, tgt = (t.to(dev) for t in make_pair((64, 64, 64), "rigid"))
    p0 = torch.tensor([0.03, -0.02, 0.04, 0.05, -0.05, 0.02], device=dev)
    single = TF.AffineProblem(mov, tgt, "rigid", p0, 5)
    single.run(5, 1e-3, 0.5, 0.5)
    paths = []
    for peer in (False, None):          # NCCL all-reduce between two kernels / fused epoch kernel with the peer-memory all-reduce
        sh = ShardedAffine(mov, tgt, "rigid", p0, 5, peer=peer)
        sh.run(2, 1e-3, 0.5, 0.5)
        sh.run(3, 1e-3, 0.5, 0.5)       # a second call continues the mailbox sequence
        assert torch.allclose(sh.losses, single.losses, rtol=2e-5), (peer, sh.losses, single.losses)
        assert torch.allclose(sh.final_theta, single.final_theta, atol=1e-6)
        thetas = [torch.empty_like(sh.final_theta) for _ in range(world)]
        dist.all_gather(thetas, sh.final_theta)
        assert all(torch.equal(t, thetas[0]) for t in thetas), "ranks diverged"      # identical redundant update
        paths.append("peer-memory" if sh.mailbox is not None else "nccl (%s)" % sh.peer_error)

    # (c) one volume, direct flow, z-slabs + halo exchange + all-reduce of 6 moments
    shape = (48, 40, 44)
    mov, tgt = (t.to(dev) for t in make_pair(shape, "flow"))
    # SGD for the tight comparison (Adam amplifies last-bit differences of the moment sums where the gradient is
    # rounding noise, see test_direct_flow_slabs_equal_whole_volume)
    whole = TF.DirectFlowProblem(mov, tgt, 4, flow0=(0.3 * smooth_flow(shape, 1.0)).to(dev), optimiser="sgd")
    whole.run(4, 0.5, 0.5, 0.5, 3.0)
    for peer in (False, None):          # NCCL halo exchange + all-reduce / neighbours' slices read in place + in-kernel all-reduce
        sd = ShardedDirectFlow(mov, tgt, 4, optimiser="sgd", peer=peer)
        sd.prob.flow.copy_((0.3 * smooth_flow(shape, 1.0)).to(dev)[:, :, sd.z0:sd.z1])
        sd.run(1, 0.5, 0.5, 0.5, 3.0)
        sd.run(3, 0.5, 0.5, 0.5, 3.0)
        assert torch.allclose(sd.flow_slab, whole.flow[:, :, sd.z0:sd.z1], atol=2e-6), (peer, (sd.flow_slab - whole.flow[:, :, sd.z0:sd.z1]).abs().max())
        assert torch.allclose(sd.losses, whole.losses, rtol=2e-5), (peer, sd.losses, whole.losses)
        paths.append("flow: " + ("peer-memory" if sd.mailbox is not None else "nccl (%s)" % sd.peer_error))

    # (a) batch of independent pairs sharded by pair, no collective in the loop
    n_pairs = 5
    pairs = [make_pair((24, 40, 56), "rigid", seed=300 + i) for i in range(n_pairs)]
    allm = torch.cat([p[0] for p in pairs]).to(dev); allt = torch.cat([p[1] for p in pairs]).to(dev)
    pp = torch.tensor([[0.01 * (i + 1), -0.01, 0.03, 0.05, -0.05, 0.02] for i in range(n_pairs)], device=dev)
    ref = TF.AffineProblem(allm, allt, "rigid", pp, 4); ref.run(4, 1e-3, 0.5, 0.5)
    a, b = shard_pairs(n_pairs, world, rank)
    out = torch.zeros(n_pairs, 3, 4, device=dev)
    if b > a:
        mine = TF.AffineProblem(allm[a:b], allt[a:b], "rigid", pp[a:b], 4); mine.run(4, 1e-3, 0.5, 0.5)
        out[a:b] = mine.final_theta
    dist.all_reduce(out)
    assert torch.allclose(out, ref.final_theta, atol=2e-6)
    dist.barrier()
    return "MGPU OK world=%d; sharded affine paths checked: %s" % (world, paths)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    msg = run_checks(rank, world, dev)
    if rank == 0:
        print(msg)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
