"""Small-shape run of the kernels added in round 2, for compute-sanitizer (memcheck / racecheck / synccheck):
persistent kernel (TMA, gather and STORE variants, 16-slice and half tiles), source-space NMI + the one-call default-loss loop
(2-D / 3-D, vector and scalar loads, batches), TMA forward warp, Edge3D, host-parameter upload."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200 as tr
import torchregister_b200.functional as TF
from torchregister_b200 import warpings as WP
from torchregister_b200.synth import make_pair
dev = "cuda"
def pairs(shape, n, kind="rigid"):
    ps = [make_pair(shape, kind, device=dev, seed=11 + i) for i in range(n)]
    return torch.cat([p[0] for p in ps]), torch.cat([p[1] for p in ps])
# persistent kernel: full + half + ragged tiles (D = 40), ragged x/y, batches, both losses, Adam, rigid rand start (gather variant)
for shape, n in (((40, 48, 64), 2), ((24, 35, 70), 3), ((16, 16, 32), 70)):
    m, t = pairs(shape, n)
    for mode, p0 in (("affine", torch.eye(3, 4).reshape(1, -1)), ("rigid", torch.tensor([[0.02, -0.01, 0.03, 0.05, -0.05, 0.02]])),
                     ("rigid", torch.tensor([[0.5, 0.77, 0.09, 0.13, 0.31, 0.63]]))):
        for w, opt in (((0.5, 0.5), "sgd"), ((1.0, 0.0), "sgd"), ((0.0, 1.0), "adam")):
            prob = TF.AffineProblem(m, t, mode, p0, 3)
            prob.run(3, 1e-4, w[0], w[1], optimiser=opt)
            assert torch.isfinite(prob.losses).all(), (shape, mode, w)
    out = TF.warp_affine(prob.final_theta, m)
    assert torch.isfinite(out).all()
# default loss through the stock call: one-call loop, 3-D (vector + scalar loads, small rotation + rand start) and 2-D
for shape, n in (((24, 32, 64), 2), ((17, 21, 33), 1), ((64, 48), 2), ((45, 51), 1)):
    m, t = pairs(shape, n)
    for mode in ("rigid", "affine"):
        torch.manual_seed(0)
        reg = tr.Register(mode=mode, device=dev)
        reg.optim(m, t, lr=1e-5, max_epochs=3)
        assert torch.isfinite(reg.losses).all(), (shape, mode)
    lo, hi = TF.NmiSourceTerm.bounds(m, t)
    term = TF.NmiSourceTerm(t, lo, hi)
    loss, g = term.loss_grad(m, 0.33)
    assert torch.isfinite(loss).all() and torch.isfinite(g).all()
# large enough for the one-pass persistent launches of the unfused passes (>= 10 tiles per SM) incl. the STORE instantiation
m, t = pairs((64, 160, 160), 3, "affine")
reg = tr.Register(mode="affine", device=dev)
reg.optim(m, t, lr=1e-5, max_epochs=2)
assert torch.isfinite(reg.losses).all()
# Edge3D
e = tr.Edge3D(device=dev)
assert torch.isfinite(e(m[:1, :, :24, :40, :40].contiguous())).all()
torch.cuda.synchronize()
print("sanitize_round2 ok")
