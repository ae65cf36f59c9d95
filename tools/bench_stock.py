"""The reference's stock call: Register(mode='rigid').optim(moving, target) — torch.rand start, DEFAULT loss (MSE+NCC+NMI)."""
import sys, time, torch
sys.path.insert(0, ".")
import torchregister_b200 as tr
from torchregister_b200.synth import make_pair
dev = "cuda:0"
for shape in ((192, 192, 160),):
    m, t = make_pair(shape, "rigid", device=dev)
    for name, kw, okw in (("stock rigid (rand start, default loss)", {}, {}),
                          ("rigid, small start, default loss", {}, {"reg0": torch.tensor([0.02, -0.01, 0.03, 0.05, -0.05, 0.02])}),
                          ("rigid, rand start, NCC only", {"weight": [0., 1., 0.]}, {}),
                          ("affine, default loss", {"mode": "affine"}, {})):
        torch.manual_seed(0)
        r = tr.Register(mode=kw.pop("mode", "rigid"), device=dev, **kw)
        r.optim(m, t, lr=1e-5, max_epochs=3, **okw)
        torch.cuda.synchronize()
        wall = []
        for ep in (40, 140):            # per-epoch slope: set-up (bounds, tables, target moments) is paid once per call
            torch.manual_seed(0)
            t0 = time.perf_counter()
            r.optim(m, t, lr=1e-5, max_epochs=ep, **okw)
            torch.cuda.synchronize()
            wall.append(time.perf_counter() - t0)
        print("%-45s %s: %.1f us/epoch (40-epoch call: %.1f)" % (name, shape, (wall[1] - wall[0]) / 100 * 1e6, wall[0] / 40 * 1e6), flush=True)
