"""Does the persistent kernel slow down in a multi-process / NCCL-initialised setting?  torchrun ... tools/mp_probe.py [nccl|none]"""
import os, sys, torch
sys.path.insert(0, ".")
mode = sys.argv[1] if len(sys.argv) > 1 else "none"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev_index = int(os.environ.get("PROBE_DEV", local))
torch.cuda.set_device(dev_index)
dev = torch.device("cuda", dev_index)
if mode == "nccl":
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair
movs, tgts = [], []
for i in range(8):
    m, t = make_pair((192, 192, 160), "affine", seed=1234 + i, device=dev)
    movs.append(m); tgts.append(t)
mov, tgt = torch.cat(movs).contiguous(), torch.cat(tgts).contiguous()
ident = torch.eye(3, 4, device=dev).reshape(1, -1)
for path in ("auto", "tma"):
    TF.set_kernel_path(path)
    prob = TF.AffineProblem(mov, tgt, "affine", ident, 30 + 600)
    prob.run(30, 1e-5, 0., 1.)
    torch.cuda.synchronize()
    if mode == "nccl":
        dist.barrier()
    res = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); prob.run(200, 1e-5, 0., 1.); b.record()
        torch.cuda.synchronize()
        res.append(a.elapsed_time(b) * 1e3 / 200)
    print("mode %s rank %d dev %d path %s: %s us/epoch | %s" % (mode, rank, dev_index, path, ["%.1f" % r for r in res], prob.lib.trb_affine_kernel_status().decode()), flush=True)
if mode == "nccl":
    dist.destroy_process_group()
