"""NMI CUDA kernels vs the PyTorch restatement (fp32 and fp64) + timing."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200.functional as TF
from torchregister_b200.utils import NMILoss, _KDEMutualInfoFn
from torchregister_b200.synth import make_pair

def torch_nmi(y, yp, dtype):
    m = NMILoss(block=4)
    yp = yp.to(dtype).clone().requires_grad_(True)
    loss = _KDEMutualInfoFn.apply(m._chunks(y.to(dtype)), m._chunks(yp), m.bins, float(m.bandwidth), float(m.alpha), m.block)
    (g,) = torch.autograd.grad(loss, yp)
    return loss.item(), g

def torch_nmi_adjoint(y, yp, dtype):
    """gradient w.r.t. the resampled values, scattered with the FORWARD's nearest indices (exact adjoint)."""
    m = NMILoss(block=4)
    nd = y.dim() - 2
    idx = [torch.clamp(torch.floor(torch.arange(200, dtype=torch.float32, device=y.device) * torch.tensor(S / 200.0, dtype=torch.float32)).long(), max=S - 1)
           for S in y.shape[2:]]
    grids = torch.meshgrid(*idx, indexing="ij")
    ws = yp.to(dtype)[(0, 0) + tuple(grids)].reshape(2 ** nd, -1).clone().requires_grad_(True)
    ts = y.to(dtype)[(0, 0) + tuple(grids)].reshape(2 ** nd, -1)
    assert torch.equal(ts, m._chunks(y.to(dtype)))
    loss = _KDEMutualInfoFn.apply(ts, ws, m.bins, float(m.bandwidth), float(m.alpha), m.block)
    (g,) = torch.autograd.grad(loss, ws)
    out = torch.zeros_like(yp.to(dtype))
    out[0, 0].index_put_(tuple(grids), g.reshape(grids[0].shape), accumulate=True)
    return loss.item(), out

for shape, scale in (((24, 32, 40), 1.0), ((24, 32, 40), 255.0), ((64, 48), 1.0), ((256, 256), 255.0), ((300, 180), 40.0),
                     ((210, 96, 230), 255.0), ((192, 192, 160), 1.0), ((192, 192, 160), 255.0)):
    mov, tgt = make_pair(shape, "rigid", device="cuda")
    y, yp = (tgt * scale).contiguous(), (mov * scale).contiguous()
    term = TF.NmiTerm(y)
    loss, g = term.loss_grad(yp, 1.0)
    torch.cuda.synchronize()
    l = loss.item()
    big = len(shape) == 3 and shape[0] > 100
    l64, g64 = torch_nmi(y, yp, torch.float64)
    l32, g32 = torch_nmi(y, yp, torch.float32)
    gs = g64.abs().max().item()
    print(shape, scale, "loss cuda %.8g t64 %.8g t32 %.8g | rel %.2e (t32 rel %.2e) | grad max %.3e cuda err %.2e t32 err %.2e"
          % (l, l64, l32, abs(l - l64) / max(abs(l64), 1e-30), abs(l32 - l64) / max(abs(l64), 1e-30), gs,
             (g.double() - g64).abs().max().item() / gs, (g32.double() - g64).abs().max().item() / gs))
    la, ga = torch_nmi_adjoint(y, yp, torch.float64)
    d = (g.double() - ga).abs()
    print("   vs exact adjoint (fp64): max err %.2e, voxels off by >1e-3*max: %d; torch autograd vs adjoint: %d voxels"
          % (d.max().item() / gs, int((d > 1e-3 * gs).sum()), int(((g64 - ga).abs() > 1e-3 * gs).sum())))
    if big:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): term.loss_grad(yp, 1.0)
        torch.cuda.synchronize(); print("   cuda ms/call %.3f" % ((time.perf_counter() - t0) / 5 * 1e3))
