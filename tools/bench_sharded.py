"""BASELINE configs[4]: ONE large volume sharded into z-slabs over the ranks (torchrun, one process per GPU).
  direct flow (halo slice exchange + fused epoch kernel + 6-value all-reduce)   and
  affine (moments kernel -> 41-value all-reduce -> apply kernel).
Per-epoch time = CUDA events around E epochs, max over ranks.  Prints one JSON line on rank 0."""
import json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    E = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from torchregister_b200.parallel import ShardedAffine, ShardedDirectFlow
    from torchregister_b200.synth import make_pair
    shape = (S, S, S)
    mov, tgt = make_pair(shape, "flow", device=dev)
    vox = S ** 3
    out = {"volume": list(shape), "n_gpus": world, "epochs": E}

    def timed(fn):
        fn(3)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(E); b.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([a.elapsed_time(b) / E], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for opt, w, lam, bpv in (("sgd", (0.5, 0.5), 2.0, 32), ("adam", (0.5, 0.5), 2.0, 80), ("sgd", (1.0, 0.0), 0.0, 32)):
        for tag, peer in (("", False), ("_fused_peer", None)):
            if peer is None and world == 1:
                continue
            sd = ShardedDirectFlow(mov, tgt, 100000, optimiser=opt, peer=peer)
            ms = timed(lambda n: sd.run(n, 0.05, w[0], w[1], lam))
            out["direct_flow_%s_mse%g_ncc%g_smooth%g%s" % (opt, w[0], w[1], lam, tag)] = {
                "ms_per_epoch": ms, "voxel_warps_per_s": vox / (ms * 1e-3), "algorithmic_GBps_total": bpv * vox / (ms * 1e-3) / 1e9,
                "path": "peer-memory" if sd.mailbox is not None else "nccl"}
            del sd
            torch.cuda.empty_cache()
    ident = torch.eye(3, 4, device=dev).reshape(1, -1)
    for name, peer in (("affine_ncc_nccl_allreduce", False), ("affine_ncc_fused_peer_allreduce", None)):
        if peer is None and world == 1:
            continue
        sa = ShardedAffine(mov, tgt, "affine", ident, 100000, peer=peer)
        ms = timed(lambda n: sa.run(n, 1e-5, 0.0, 1.0, align=False))
        out[name] = {"ms_per_epoch": ms, "voxel_warps_per_s": vox / (ms * 1e-3), "algorithmic_GBps_total": 8 * vox / (ms * 1e-3) / 1e9,
                     "path": "peer-memory" if sa.mailbox is not None else "nccl", "first_losses": sa.losses[0, :2].tolist()}
        del sa
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
