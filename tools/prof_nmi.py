import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
mov, tgt = make_pair((192, 192, 160), "rigid", device="cuda")
term = TF.NmiTerm((tgt * scale).contiguous())
for _ in range(3):
    term.loss_grad((mov * scale).contiguous(), 1.0)
torch.cuda.synchronize()
