"""Where an e2e step (bench.py) spends its time: wall clock of each public-API call on a resident batch."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200 as tr
from torchregister_b200.synth import make_pair
dev = "cuda"
pairs = [make_pair((192, 192, 160), "affine", seed=1234 + i, device=dev) for i in range(8)]
m = torch.cat([p[0] for p in pairs]); t = torch.cat([p[1] for p in pairs])
reg0 = torch.tensor([0.02, -0.01, 0.03, 0.05, -0.05, 0.02]).repeat(8, 1)
def wall(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return out, (t1 - t0) * 1e3, (t2 - t0) * 1e3
for rep in range(2):
    r = tr.Register(mode="rigid", device=dev, weight=[0.0, 1.0, 0.0])
    _, host, tot = wall(lambda: r.optim(m, t, lr=1e-5, max_epochs=500, reg0=reg0)); print("rigid.optim(500): host %.1f ms, done %.1f ms (kernel-only %.1f)" % (host, tot, 500 * 0.158))
    m2, host, tot = wall(lambda: r(m)); print("warp: host %.1f ms, done %.1f ms" % (host, tot))
    a = tr.Register(mode="affine", device=dev, weight=[0.0, 1.0, 0.0])
    _, host, tot = wall(lambda: a.optim(m2, t, lr=1e-5, max_epochs=200)); print("affine.optim(200): host %.1f ms, done %.1f ms (kernel-only %.1f)" % (host, tot, 200 * 0.158))
