"""Epoch time of the batched TMA kernel by mode / starting point (8 pairs x 192x192x160, NCC)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair
dev = "cuda"
pairs = [make_pair((192, 192, 160), "affine", seed=1234 + i, device=dev) for i in range(8)]
mov = torch.cat([p[0] for p in pairs]); tgt = torch.cat([p[1] for p in pairs])
def timed(prob, n, *a):
    prob.run(20, *a); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prob.run(n, *a); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
ident = torch.eye(3, 4, device=dev).reshape(1, -1)
print("affine from identity        %.1f us" % timed(TF.AffineProblem(mov, tgt, "affine", ident, 400), 200, 1e-5, 0.0, 1.0))
th = torch.tensor([[1.01, .02, -.01, .03, -.02, .99, .01, -.02, .01, -.01, 1., .01]], device=dev)
print("affine from a rotated theta %.1f us" % timed(TF.AffineProblem(mov, tgt, "affine", th, 400), 200, 1e-5, 0.0, 1.0))
reg0 = torch.tensor([0.02, -0.01, 0.03, 0.05, -0.05, 0.02], device=dev).repeat(8, 1)
print("rigid from reg0             %.1f us" % timed(TF.AffineProblem(mov, tgt, "rigid", reg0, 400), 200, 1e-5, 0.0, 1.0))
print("rigid from zeros            %.1f us" % timed(TF.AffineProblem(mov, tgt, "rigid", torch.zeros(8, 6, device=dev), 400), 200, 1e-5, 0.0, 1.0))
