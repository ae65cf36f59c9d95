"""BASELINE configs[0] as the stock call makes it: 2-D rigid 256x256 with the reference's DEFAULT loss (MSE+NCC+NMI):
per-epoch slope of Register.optim (wall clock), one-call loop (source-space NMI) vs the per-epoch loop."""
import sys, time, torch
sys.path.insert(0, ".")
import torchregister_b200 as tr
from torchregister_b200 import warpings as WP
from torchregister_b200.synth import make_pair
m, t = make_pair((256, 256), "rigid", device="cuda:0")
p0 = torch.tensor([0.03, 0.02, -0.01])
for form in ("resampled", "source"):
    WP.set_nmi_form(form)
    reg = tr.Register(mode="rigid", device="cuda:0")
    reg.optim(m, t, lr=1e-5, max_epochs=5, reg0=p0)
    wall = []
    for ep in (100, 500):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        reg.optim(m, t, lr=1e-5, max_epochs=ep, reg0=p0)
        torch.cuda.synchronize(); wall.append(time.perf_counter() - t0)
    print("%-9s %.1f us/epoch (slope), 500 epochs in %.1f ms" % (form, (wall[1] - wall[0]) / 400 * 1e6, wall[1] * 1e3), flush=True)
WP.set_nmi_form("auto")
