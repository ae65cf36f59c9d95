"""Time of one affine epoch with the reference's DEFAULT loss (0.33 MSE + 0.33 NCC + 0.33 NMI) at 192x192x160."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200 as tr
from torchregister_b200.synth import make_pair
shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (192, 192, 160)
mov, tgt = make_pair(shape, "rigid", device="cuda")
for w, ep in (([0.33, 0.33, 0.33], 6), ([0.5, 0.5, 0.0], 200)):
    r = tr.Register(mode="affine", device="cuda", weight=w)
    r.optim(mov, tgt, lr=1e-5, max_epochs=2)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r.optim(mov, tgt, lr=1e-5, max_epochs=ep)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("weights", w, "ms/epoch %.3f" % (dt / ep * 1e3), "losses", [float(x) for x in r.losses[:2]])
