"""One fused direct-flow epoch per variant at 256^3 (for ncu)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair, smooth_flow
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shape = (S, S, S)
mov, tgt = make_pair(shape, "flow", device="cuda")
flow0 = smooth_flow(shape, 3.0, device="cuda")
for opt in ("sgd", "adam"):
    for w, sm in (((0.5, 0.5), 2.0), ((1.0, 0.0), 0.0)):
        prob = TF.DirectFlowProblem(mov, tgt, 100, flow0=flow0, optimiser=opt)
        prob.run(3, 0.05, w[0], w[1], sm)
torch.cuda.synchronize()
