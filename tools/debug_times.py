"""Per-CTA timeline of one launch of the TMA kernel (needs a TRB_TIMING=1 build)."""
import sys, os, ctypes, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200.functional as TF
from torchregister_b200 import _lib
from torchregister_b200.synth import make_pair
lib = _lib.load()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
SHAPE = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (192, 192, 160)
movs, tgts = [], []
for i in range(P):
    m, t = make_pair(SHAPE, "affine", seed=1234+i, device="cuda")
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
tiles = a[:, 8]; waits = a[:, 7] / 1e3; pubs = a[:, 9] / 1e3
print("warp 0: tiles per CTA %.1f; time in mbar_wait per CTA: mean %.1f us max %.1f us (%.2f us per tile); publishes per CTA %.1f, %.1f us each; main loop per tile %.2f us"
      % (tiles.mean(), waits.mean(), waits.max(), waits.mean() / max(tiles.mean(), 1), a[:, 10].mean(), pubs.sum() / max(a[:, 10].sum(), 1), np.median(end) / max(tiles.mean(), 1)))
n2 = 4096 + 4 * 300
buf2 = (ctypes.c_ulonglong * n2)()
lib.trb_debug_read(buf2, n2)
tl = np.array(buf2, dtype=np.uint64)[4096:].reshape(-1, 4).astype(np.int64)
tl = tl[tl[:, 0] > 0]
if len(tl):
    base = tl[0, 0]
    arrive0, go0, arrive15, go15 = [(tl[:, i] - base) / 1e3 for i in range(4)]
    d = np.diff(go0)
    print("CTA 5 warp 0: %d tiles; tile-to-tile us:" % len(tl), " ".join("%.2f" % v for v in d[:80]))
    print("   wait at barrier (warp0) us:", " ".join("%.2f" % v for v in (go0 - arrive0)[:80]))
    print("   warp15 - warp0 arrival skew us:", " ".join("%.2f" % v for v in (arrive15 - arrive0)[:80]))
