"""Quick timing of the persistent kernel: batch 8x192x192x160 (NCC, MSE), 256^3, single pair."""
import sys, torch
sys.path.insert(0, ".")
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair
dev = torch.device("cuda:0")
def timeit(shape, n_pairs, epochs, w=(0.0, 1.0), mode="affine"):
    movs, tgts = [], []
    for i in range(n_pairs):
        m, t = make_pair(shape, "affine", seed=1234 + i, device=dev)
        movs.append(m); tgts.append(t)
    mov, tgt = torch.cat(movs).contiguous(), torch.cat(tgts).contiguous()
    if mode == "affine":
        p0 = torch.eye(3, 4, device=dev).reshape(1, -1)
    else:
        p0 = torch.tensor([0.02, -0.01, 0.03, 0.05, -0.05, 0.02], device=dev).repeat(n_pairs, 1)
    prob = TF.AffineProblem(mov, tgt, mode, p0, 30 + 3 * epochs)
    prob.run(30, 1e-5, *w)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); prob.run(epochs, 1e-5, *w); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / epochs)
    return best
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for name, shape, n, ep, w, mode in (("batch8 NCC", (192, 192, 160), 8, 200, (0., 1.), "affine"), ("batch8 MSE", (192, 192, 160), 8, 200, (1., 0.), "affine"),
                                    ("256^3 NCC", (256, 256, 256), 1, 200, (0., 1.), "affine"), ("256^3 rigid NCC", (256, 256, 256), 1, 200, (0., 1.), "rigid"),
                                    ("single 192x192x160", (192, 192, 160), 1, 200, (0., 1.), "affine"), ("512^3 NCC", (512, 512, 512), 1, 40, (0., 1.), "affine")):
    us = timeit(shape, n, ep, w, mode)
    vox = shape[0] * shape[1] * shape[2] * n
    print("%s %-20s %.1f us/epoch  frac %.3f" % (tag, name, us, 8.0 * vox / us / 1e3 / 6549.8), flush=True)
