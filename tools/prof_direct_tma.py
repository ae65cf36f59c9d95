import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair, smooth_flow
shape = (256, 256, 256)
mov, tgt = make_pair(shape, "flow", device="cuda")
flow0 = smooth_flow(shape, 3.0, device="cuda")
for w in ((1.0, 0.0), (0.5, 0.5)):
    prob = TF.DirectFlowProblem(mov, tgt, 100, flow0=flow0, optimiser="sgd")
    prob.run(3, 0.05, w[0], w[1], 2.0)
torch.cuda.synchronize()
