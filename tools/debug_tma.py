import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair
shape = tuple(int(a) for a in (sys.argv[1:4] or (40, 36, 44)))
mov, tgt = make_pair(shape, "rigid")
p0 = torch.tensor([0.05, -0.03, 0.04, 0.1, -0.08, 0.05])
for path in ("direct", "auto"):
    TF.set_kernel_path(path)
    prob = TF.AffineProblem(mov.cuda(), tgt.cuda(), "rigid", p0.cuda(), 3)
    prob.run(3, 1e-3, 0.5, 0.5)
    torch.cuda.synchronize()
    print(path, prob.losses.cpu().numpy(), prob.final_theta.cpu().numpy().ravel()[:4])
