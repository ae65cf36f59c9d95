"""Per-CTA timeline of one launch of the TMA kernel (needs a TRB_TIMING=1 build)."""
import sys, os, ctypes, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200.functional as TF
from torchregister_b200 import _lib
from torchregister_b200.synth import make_pair
lib = _lib.load()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
movs, tgts = [], []
for i in range(P):
    m, t = make_pair((192,192,160), "affine", seed=1234+i, device="cuda")
    movs.append(m); tgts.append(t)
mov = torch.cat(movs); tgt = torch.cat(tgts)
prob = TF.AffineProblem(mov, tgt, "affine", torch.eye(3,4,device="cuda").reshape(1,-1), 20)
prob.run(10, 1e-5, 0., 1.); torch.cuda.synchronize()
lib.trb_debug_clear(); torch.cuda.synchronize()
prob.run(1, 1e-5, 0., 1.); torch.cuda.synchronize()
n = 148*16
buf = (ctypes.c_ulonglong * n)()
lib.trb_debug_read(buf, n)
a = np.array(buf, dtype=np.uint64).reshape(148, 16).astype(np.int64)
t0 = a[:,0].min()
end = (a[:,1]-t0)/1e3
print("pairs", P, "main-loop end us: min %.1f med %.1f max %.1f" % (end.min(), np.median(end), end.max()))
fin = a[:,2].max()
for name, c in (("is_last known", 3), ("after fence", 4), ("slots summed", 5), ("epilogues done", 6), ("zeroed/exit", 2)):
    print("   %-16s %.1f us" % (name, (a[:, c].max() - t0) / 1e3))
print("final phase done at %.1f us (tail %.1f us after the last CTA's main loop)" % ((fin-t0)/1e3, (fin-t0)/1e3 - end.max()))
