"""Short run of the persistent kernel for ncu: python tools/prof_persist.py [pairs] [epochs]"""
import sys, torch
sys.path.insert(0, ".")
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
E = int(sys.argv[2]) if len(sys.argv) > 2 else 6
SHAPE = tuple(int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (192, 192, 160)
dev = torch.device("cuda:0")
movs, tgts = [], []
for i in range(P):
    m, t = make_pair(SHAPE, "affine", seed=1234 + i, device=dev)
    movs.append(m); tgts.append(t)
mov, tgt = torch.cat(movs).contiguous(), torch.cat(tgts).contiguous()
prob = TF.AffineProblem(mov, tgt, "affine", torch.eye(3, 4, device=dev).reshape(1, -1), 3 * E)
for _ in range(3):
    prob.run(E, 1e-5, 0.0, 1.0)
torch.cuda.synchronize()
print("done", prob.losses[0, :3].tolist())
