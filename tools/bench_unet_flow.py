"""BASELINE configs[2] as the reference runs it: Register(mode='flow') = Attention_UNet (PyTorch/cuDNN) -> flow -> fused
warp + similarity + gradient node, 256^3, a few epochs; U-Net / node split."""
import sys, os, json, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn as nn
import torchregister_b200 as tr
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
E = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = "cuda:0"
mov, tgt = make_pair((S, S, S), "flow", device=dev)
torch.manual_seed(0)
out = {"size": S, "epochs": E}
for name, crit, w in (("mse+ncc", [nn.MSELoss(), tr.NCCLoss()], [0.5, 0.5]), ("mse", [nn.MSELoss()], [1.0])):
    fr = tr.flow_register((S, S, S), mode="bilinear", n=32, lr=1e-3, max_epochs=1, criterions=crit, weights=w, stop_crit=-1.0).to(dev)
    fr.optimize(mov, tgt, dev, debug=False)           # warm-up epoch (cuDNN autotune)
    torch.cuda.synchronize()
    fr.max_epochs = E
    torch.cuda.reset_peak_memory_stats()
    t0 = time.perf_counter()
    fr.optimize(mov, tgt, dev, debug=False)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / E * 1e3
    flow = fr.flow.detach()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        TF.flow_loss_grad(mov, tgt, flow, w[0], w[1] if len(w) > 1 else 0.0)
    b.record(); torch.cuda.synchronize()
    node_ms = a.elapsed_time(b) / 10
    out[name] = {"ms_per_epoch": ms, "node_ms": node_ms, "unet_and_optimizer_ms": ms - node_ms, "node_share": node_ms / ms,
                 "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9, "voxel_warps_per_s": S ** 3 / (ms * 1e-3), "losses": fr.losses[:E]}
    del fr
    torch.cuda.empty_cache()
print(json.dumps(out))
