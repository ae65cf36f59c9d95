import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair
S=256
mov, tgt = make_pair((S,S,S), "flow", device="cuda")
th = torch.tensor([[1.01, .02, -.01, .01], [-.02, .99, .01, 0.], [.01, -.01, 1., .02]], device="cuda")
def timeit(fn, n=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n
print("warp_affine us", timeit(lambda: TF.warp_affine(th, mov)))
mov2 = torch.cat([mov, mov], 1)
print("warp_affine 2ch us", timeit(lambda: TF.warp_affine(th, mov2)))
