import sys, torch
sys.path.insert(0, ".")
import torchregister_b200 as tr
from torchregister_b200.synth import make_pair
dev = "cuda:0"
m, t = make_pair((192, 192, 160), "affine", device=dev)
rd = tr.Register(mode="affine", device=dev)
rd.optim(m, t, lr=1e-5, max_epochs=int(sys.argv[1]) if len(sys.argv) > 1 else 3)
torch.cuda.synchronize()
print("ok", rd.losses[:3].tolist())
