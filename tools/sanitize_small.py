"""Small-shape run of the kernels added late in round 1, for compute-sanitizer (memcheck / synccheck / racecheck)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair, smooth_flow
dev = "cuda"
# NMI: moment form ([0,1] data), grouped form (x255), direct form (x65535); 3-D up-sampling, 2-D down-sampling
for shape, scale in (((12, 16, 20), 1.0), ((12, 16, 20), 255.0), ((12, 16, 20), 65535.0), ((230, 210), 255.0), ((230, 210), 1.0)):
    mov, tgt = make_pair(shape, "rigid", device=dev)
    term = TF.NmiTerm((tgt * scale).contiguous())
    loss, g = term.loss_grad((mov * scale).contiguous(), 1.0)
    assert torch.isfinite(loss).all() and torch.isfinite(g).all()
# fused direct flow: register-staged (W % 4 != 0), TMA-staged (W = 40), all optimiser / loss variants, slab with halos
for shape in ((9, 12, 37), (9, 12, 40), (11, 19, 72)):
    mov, tgt = make_pair(shape, "flow", device=dev)
    f0 = (0.3 * smooth_flow(shape, 1.0)).to(dev)
    for opt in ("sgd", "adam"):
        for w, sm in (((1.0, 0.0), 0.0), ((1.0, 0.0), 3.0), ((0.5, 0.5), 3.0), ((0.5, 0.5), 0.0)):
            prob = TF.DirectFlowProblem(mov, tgt, 4, flow0=f0, optimiser=opt)
            prob.run(3, 0.1, w[0], w[1], sm)
            assert torch.isfinite(prob.flow).all() and torch.isfinite(prob.losses).all()
    a, b = 3, 7
    slab = TF.DirectFlowProblem(mov, tgt[:, :, a:b].contiguous(), 3, z_off=a, flow0=f0[:, :, a:b].contiguous())
    lo, hi = f0[0, :, a - 1].contiguous(), f0[0, :, b].contiguous()
    slab.prime(0.5)
    slab.step(0.1, 0.5, 0.5, 3.0, lo, hi)
    slab.finish(0.5, 0.5, 3.0)
    assert torch.isfinite(slab.flow).all()
torch.cuda.synchronize()
print("sanitize_small ok")
