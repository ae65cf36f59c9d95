"""Source-space NMI term (csrc/nmi_src.cu) against the resampled-array form (csrc/nmi.cu), and the one-call default-loss
loop against the per-epoch loop: values, gradients, trajectories, time per epoch."""
import sys, time, torch
sys.path.insert(0, ".")
import torchregister_b200 as tr
from torchregister_b200 import functional as TF, warpings as WP
from torchregister_b200.synth import make_pair

dev = "cuda:0"
ok = True
for shape, n in [((24, 32, 64), 2), ((40, 48, 36), 1), ((160, 192, 192), 1), ((210, 96, 230), 2), ((256, 256, 256), 1)]:
    torch.manual_seed(3)
    ms, ts = zip(*[make_pair(shape, "affine", device=dev, seed=i) for i in range(n)])
    m, t = torch.cat(ms), torch.cat(ts)
    lo, hi = TF.NmiSourceTerm.bounds(m, t)
    src = TF.NmiSourceTerm(t, lo, hi)
    ls, gs = src.loss_grad(m, 0.33)
    worst_l = worst_g = 0.0
    for i in range(n):
        old = TF.NmiTerm(t[i:i + 1])
        lo_, go_ = old.loss_grad(m[i:i + 1], 0.33)
        worst_l = max(worst_l, abs(ls[i].item() - lo_.item()) / abs(lo_.item()))
        worst_g = max(worst_g, ((gs[i:i + 1] - go_).abs().max() / go_.abs().max()).item())
    good = worst_l < 1e-5 and worst_g < 1e-4
    ok &= good
    print("term %s x%d bounds [%.3f, %.3f] loss %.6g  rel loss diff %.2e  grad diff / max %.2e %s"
          % (shape, n, lo, hi, ls[0].item(), worst_l, worst_g, "OK" if good else "FAIL"), flush=True)

for shape, n, mode, ep in [((24, 32, 64), 2, "rigid", 8), ((160, 192, 192), 1, "affine", 6)]:
    ms, ts = zip(*[make_pair(shape, mode, device=dev, seed=i) for i in range(n)])
    m, t = torch.cat(ms), torch.cat(ts)
    res = {}
    for form in ("resampled", "source"):
        WP.set_nmi_form(form)
        nd = 3
        p0 = torch.zeros(1, 6) if mode == "rigid" else torch.eye(3, 4).reshape(1, -1)
        prob, _, (ft, bt) = WP._affine_like(mode, m, t, 1e-5 if mode == "affine" else 1e-3, ep, (0.33, 0.33, 0.33), p0, False, want_warped=False)
        res[form] = (prob.losses.clone(), ft.clone())
    WP.set_nmi_form("auto")
    rl = ((res["source"][0] - res["resampled"][0]).abs() / res["resampled"][0].abs()).max().item()
    rt = (res["source"][1] - res["resampled"][1]).abs().max().item()
    good = rl < 2e-5 and rt < 2e-6
    ok &= good
    print("loop %s x%d %s: loss rel %.2e theta %.2e  losses %s %s" % (shape, n, mode, rl, rt, res["source"][0][0, :3].tolist(), "OK" if good else "FAIL"), flush=True)

for shape, n in [((160, 192, 192), 1), ((256, 256, 256), 1), ((160, 192, 192), 8)]:
    ms, ts = zip(*[make_pair(shape, "affine", device=dev, seed=i) for i in range(n)])
    m, t = torch.cat(ms), torch.cat(ts)
    for form, ep in (("resampled", 10), ("source", 50)):
        WP.set_nmi_form(form)
        p0 = torch.eye(3, 4).reshape(1, -1)
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            WP._affine_like("affine", m, t, 1e-5, ep, (0.33, 0.33, 0.33), p0, False, want_warped=False)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print("time %s x%d %-9s %.1f us/epoch (incl. setup, %d epochs)" % (shape, n, form, dt / ep * 1e6, ep), flush=True)
    WP.set_nmi_form("auto")
print("NMI SRC", "OK" if ok else "FAIL")
