"""Persistent multi-epoch kernel vs the per-epoch TMA kernel and the direct kernel: same losses / theta, and timings."""
import sys, json, time
import torch
sys.path.insert(0, ".")
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair

dev = torch.device("cuda:0")

def run(path, shape, n_pairs, mode, epochs, w, lr, opt="sgd"):
    TF.set_kernel_path(path)
    movs, tgts = [], []
    for i in range(n_pairs):
        m, t = make_pair(shape, mode if mode != "rigid" else "rigid", seed=100 + i, device=dev)
        movs.append(m); tgts.append(t)
    mov, tgt = torch.cat(movs).contiguous(), torch.cat(tgts).contiguous()
    if mode == "rigid":
        p0 = torch.tensor([0.02, -0.01, 0.03, 0.05, -0.05, 0.02], device=dev).repeat(n_pairs, 1)
    else:
        p0 = torch.eye(3, 4, device=dev).reshape(1, -1)
    prob = TF.AffineProblem(mov, tgt, mode, p0, epochs)
    prob.run(epochs, lr, w[0], w[1], optimiser=opt)
    torch.cuda.synchronize()
    return prob.losses.cpu().double(), prob.final_theta.cpu().double(), prob.best_theta.cpu().double()

ok = True
cases = [((40, 48, 64), 1, "rigid", 12, (0.5, 0.5), 1e-3, "sgd"),
         ((40, 48, 64), 3, "affine", 12, (0.0, 1.0), 1e-4, "sgd"),
         ((24, 32, 64), 5, "rigid", 9, (1.0, 0.0), 5e-2, "sgd"),
         ((37, 50, 36), 2, "affine", 7, (0.3, 0.7), 1e-4, "adam"),
         ((64, 64, 96), 20, "rigid", 5, (0.5, 0.5), 1e-3, "sgd"),
         ((96, 96, 96), 1, "affine", 300, (0.0, 1.0), 1e-4, "sgd")]
for c in cases:
    a = run("auto", *c)
    b = run("tma", *c)
    d = run("direct", *c)
    rl = ((a[0] - b[0]).abs() / b[0].abs().clamp_min(1e-30)).max().item()
    rt = (a[1] - b[1]).abs().max().item()
    rb = (a[2] - b[2]).abs().max().item()
    rl2 = ((a[0] - d[0]).abs() / d[0].abs().clamp_min(1e-30)).max().item()
    rt2 = (a[1] - d[1]).abs().max().item()
    a2 = run("auto", *c)
    rep = bool((a2[0] == a[0]).all() and (a2[1] == a[1]).all())
    good = rl < 5e-5 and rt < 2e-6 and rb < 2e-6 and rep      # north-star tolerance: 1e-4 / 1e-4
    ok &= good
    print("case", c[:4], c[6], "loss rel vs tma %.2e theta %.2e best %.2e | vs direct %.2e %.2e | reproducible %s %s"
          % (rl, rt, rb, rl2, rt2, rep, "OK" if good else "FAIL"), flush=True)
print("PARITY", "OK" if ok else "FAIL", flush=True)

def timeit(path, shape, n_pairs, epochs, w=(0.0, 1.0)):
    TF.set_kernel_path(path)
    movs, tgts = [], []
    for i in range(n_pairs):
        m, t = make_pair(shape, "affine", seed=1234 + i, device=dev)
        movs.append(m); tgts.append(t)
    mov, tgt = torch.cat(movs).contiguous(), torch.cat(tgts).contiguous()
    ident = torch.eye(3, 4, device=dev).reshape(1, -1)
    prob = TF.AffineProblem(mov, tgt, "affine", ident, 30 + epochs)
    prob.run(30, 1e-5, *w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prob.run(epochs, 1e-5, *w); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / epochs

res = {}
for name, shape, n, ep in (("batch8_192x192x160", (192, 192, 160), 8, 200), ("single_192x192x160", (192, 192, 160), 1, 200),
                           ("single_256^3", (256, 256, 256), 1, 200), ("single_512^3", (512, 512, 512), 1, 40)):
    for path in ("auto", "tma"):
        us = timeit(path, shape, n, ep)
        vox = shape[0] * shape[1] * shape[2] * n
        res[name + "_" + path] = {"us_per_epoch": us, "GBps": 8.0 * vox / us / 1e3, "frac": 8.0 * vox / us / 1e3 / 6549.8}
        print(name, path, "%.1f us/epoch  %.0f GB/s  frac %.3f" % (us, 8.0 * vox / us / 1e3, 8.0 * vox / us / 1e3 / 6549.8), flush=True)
us = timeit("auto", (192, 192, 160), 8, 200, (1.0, 0.0))
print("batch8 mse-only auto %.1f us" % us)
res["batch8_mse_auto"] = us
json.dump(res, open("gpurun_out/persist_check.json", "w"), indent=1)
