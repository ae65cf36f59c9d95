"""BASELINE configs[0]: 2-D rigid 256x256, 500 epochs — our GPU path vs the CPU oracle port."""
import sys, os, time, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200 as tr
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair
mov, tgt = make_pair((256, 256), "rigid")
m, t = mov.cuda(), tgt.cuda()
p0 = torch.tensor([0.03, 0.02, -0.01], device="cuda")
out = {}
for name, w in (("mse_only", (1.0, 0.0)), ("mse+ncc", (0.5, 0.5))):
    prob = TF.AffineProblem(m, t, "rigid", p0, 1000)
    prob.run(100, 1e-5, *w); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); prob.run(500, 1e-5, *w); b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / 500
    out[name] = {"us_per_epoch": us, "iters_per_s": 1e6 / us, "voxel_warps_per_s": 65536 / (us * 1e-6)}
t0 = time.perf_counter()
reg = tr.Register(mode="rigid", device="cuda", weight=[0.5, 0.5, 0.0])
reg.optim(mov, tgt, lr=1e-5, max_epochs=500, reg0=p0.cpu())
th = reg.theta.cpu(); dt = time.perf_counter() - t0
out["register_api_500_epochs_wall_s"] = dt
print(json.dumps(out))
