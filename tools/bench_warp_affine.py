"""Forward affine warp (Register.__call__ / get_affine_warp): TMA-staged kernel vs the gather kernel, parity and timing."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair
dev = "cuda"
def timeit(fn, n=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n
th = torch.tensor([[1.01, .02, -.01, .01], [-.02, .99, .01, 0.], [.01, -.01, 1., .02]], device=dev)
big = torch.tensor([[0.7, -0.7, 0.1, .05], [0.7, 0.7, -0.1, 0.], [0.0, 0.14, 0.99, .02]], device=dev)
ok = True
for shape in ((40, 48, 64), (37, 50, 36), (24, 32, 128)):
    m, _ = make_pair(shape, "flow", device=dev)
    batch = torch.cat([torch.cat([m, 0.5 * m + 0.1], 1), torch.cat([0.3 * m, m * m], 1), torch.cat([m, m], 1)], 0).contiguous()   # 3 pairs x 2 channels
    ths = torch.stack([th, big, th * 1.0])
    a = TF.warp_affine(ths, batch)
    g = TF.warp_affine(ths, batch, large_rotation=True)
    err = (a - g).abs().max().item()
    ok &= err < 1e-5
    print(shape, "tma vs gather kernel max abs diff %.2e" % err)
print("WARP PARITY", "OK" if ok else "FAIL")
for S in (256,):
    mov, tgt = make_pair((S, S, S), "flow", device=dev)
    vol = S ** 3
    for name, t in (("small", th), ("45deg", big)):
        for large in (False, True):
            us = timeit(lambda: TF.warp_affine(t, mov, large_rotation=large))
            print("%d^3 %s %s: %.1f us  %.0f GB/s (8 B/voxel) frac %.3f" % (S, name, "gather" if large else "tma", us, 8 * vol / us / 1e3, 8 * vol / us / 1e3 / 6549.8))
    mov2 = torch.cat([mov, mov], 1)
    print("2 channels tma us", timeit(lambda: TF.warp_affine(th, mov2)))
m8 = torch.cat([make_pair((192, 192, 160), "affine", seed=i, device=dev)[0] for i in range(8)])
th8 = th.repeat(8, 1, 1)
us = timeit(lambda: TF.warp_affine(th8, m8))
print("batch 8 x 192x192x160 one launch: %.1f us  %.0f GB/s" % (us, 8 * m8.numel() / us / 1e3))
