"""One default-loss (MSE + NCC + NMI) affine loop, for `ncu --metrics gpu__time_duration.sum` launch lists and timing:
python tools/prof_default_loss.py [epochs] [D H W] [pairs] [form]"""
import sys, time, torch
sys.path.insert(0, ".")
import torchregister_b200 as tr
from torchregister_b200 import warpings as WP
from torchregister_b200.synth import make_pair
dev = "cuda:0"
ep = int(sys.argv[1]) if len(sys.argv) > 1 else 3
shape = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (160, 192, 192)
n = int(sys.argv[5]) if len(sys.argv) > 5 else 1
WP.set_nmi_form(sys.argv[6] if len(sys.argv) > 6 else "auto")
ms, ts = zip(*[make_pair(shape, "affine", device=dev, seed=i) for i in range(n)])
m, t = torch.cat(ms), torch.cat(ts)
p0 = torch.eye(3, 4).reshape(1, -1)
def wall(e):
    best = 1e9
    for _ in range(3):              # min of 3: the set-up of a call (allocator, bounds, target moments) varies by milliseconds
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        prob, _, _ = WP._affine_like("affine", m, t, 1e-5, e, (0.33, 0.33, 0.33), p0, False, want_warped=False)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, prob
wall(2)
out = [0.0, wall(ep)[0]]
t5, prob = wall(5 * ep)
out.append(t5)
print("shape %s x%d: %.1f us/epoch (difference of a %d- and a %d-epoch call)" % (shape, n, (out[2] - out[1]) / (4 * ep) * 1e6, ep, 5 * ep))
print("losses", prob.losses[0, :3].tolist())
