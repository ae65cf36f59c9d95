"""The stock flow call: Register(mode='flow').optim(moving, target) with the reference's defaults (U-Net n = 32, MSE + NCC + NMI)."""
import sys, time, torch
sys.path.insert(0, ".")
import torchregister_b200 as tr
from torchregister_b200.synth import make_pair
dev = "cuda:0"
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m, t = make_pair((S, S, S), "flow", device=dev)
for kw in ({}, {"weight": [0.5, 0.5, 0.0]}):
    wall = []
    for ep in (2, 3, 9):
        torch.manual_seed(0)
        r = tr.Register(mode="flow", device=dev, **kw)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r.optim(m, t, lr=1e-3, max_epochs=ep)
        torch.cuda.synchronize(); wall.append(time.perf_counter() - t0)
    print("Register(mode='flow', %s) %d^3: %.1f ms/epoch (slope), losses %s" % (kw, S, (wall[2] - wall[1]) / 6 * 1e3, [round(float(v), 5) for v in r.losses[:3]]), flush=True)
