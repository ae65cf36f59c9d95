"""Where a persistent-kernel CTA spends its time (needs a TRB_TIMING=1 build): python tools/debug_persist.py P D H W EPOCHS"""
import sys, os, ctypes, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200.functional as TF
from torchregister_b200 import _lib
from torchregister_b200.synth import make_pair
lib = _lib.load()
lib.trb_pdebug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
P = int(sys.argv[1]); SHAPE = tuple(int(v) for v in sys.argv[2:5]); E = int(sys.argv[5])
movs, tgts = [], []
for i in range(P):
    m, t = make_pair(SHAPE, "affine", seed=1234 + i, device="cuda")
    movs.append(m); tgts.append(t)
mov = torch.cat(movs); tgt = torch.cat(tgts)
prob = TF.AffineProblem(mov, tgt, "affine", torch.eye(3, 4, device="cuda").reshape(1, -1), 10 + E)
prob.run(10, 1e-5, 0., 1.); torch.cuda.synchronize()
lib.trb_pdebug_clear(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); prob.run(E, 1e-5, 0., 1.); e1.record(); torch.cuda.synchronize()
n = 8192 + 1200
buf = (ctypes.c_ulonglong * n)()
lib.trb_pdebug_read(buf, n)
raw = np.array(buf, dtype=np.uint64).astype(np.int64)
a = raw[:148 * 32].reshape(148, 32)
print("pairs", P, SHAPE, "epochs", E, "event us/epoch %.1f" % (e0.elapsed_time(e1) * 1e3 / E))
def us(c): return a[:, c].mean() / 1e3 / E
print("per epoch per CTA (mean over CTAs), us:")
print("  producer: acquire total %.2f (count %.1f; poll %.2f [retries %.1f], epilogue %.2f), new-column work %.2f, wait empty %.2f; stream %.1f"
      % (us(0), a[:, 1].mean() / E, us(2), a[:, 16].mean() / E, us(3), us(5), us(4), (a[:, 7] - a[:, 6]).mean() / 1e3 / E))
print("  consumer w0: wait full %.2f (w15 %.2f), tiles %.1f, column end %.2f (cols %.1f), column setup %.2f"
      % (us(8), us(13), a[:, 9].mean() / E, us(10), a[:, 11].mean() / E, us(12)))
print("  reducer: wait %.2f, work %.2f" % (us(14), us(15)))
tl = raw[8192:].reshape(-1, 2)
tl = tl[tl[:, 0] > 0]
if len(tl) > 2:
    go = (tl[:, 1] - tl[0, 0]) / 1e3
    d = np.diff(go)
    print("CTA 5 warp 0 tile-to-tile us:", " ".join("%.2f" % v for v in d[:120]))
    print("   wait at full barrier us:", " ".join("%.2f" % v for v in ((tl[:, 1] - tl[:, 0]) / 1e3)[:120]))
