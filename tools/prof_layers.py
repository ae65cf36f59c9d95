"""One flow-mode epoch at 256^3 and a few default-loss epochs, for ncu metric passes over the layer kernels
(instnorm / thinconv / pointconv / nmi_src): python tools/prof_layers.py"""
import sys, torch
sys.path.insert(0, ".")
import torch.nn as nn
import torchregister_b200 as tr
from torchregister_b200.synth import make_pair
dev = "cuda:0"
m, t = make_pair((256, 256, 256), "flow", device=dev)
torch.manual_seed(0)
fr = tr.flow_register((256, 256, 256), mode="bilinear", n=32, lr=1e-3, max_epochs=2, criterions=[nn.MSELoss(), tr.NCCLoss()],
                      weights=[0.5, 0.5], stop_crit=-1.0).to(dev)
fr.optimize(m, t, dev, debug=False)
m, t = make_pair((192, 192, 160), "affine", device=dev)
r = tr.Register(mode="affine", device=dev)
r.optim(m, t, lr=1e-5, max_epochs=2)
torch.cuda.synchronize()
print("done")
