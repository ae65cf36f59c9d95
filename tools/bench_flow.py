"""Timing of the flow-side kernels at BASELINE configs[2] size (256^3): fused flow node, direct-flow epoch, warps."""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200.functional as TF
from torchregister_b200.synth import make_pair, smooth_flow
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shape = (S, S, S)
dev = "cuda"
mov, tgt = make_pair(shape, "flow", device=dev)
flow = smooth_flow(shape, 3.0, device=dev)
vox = S ** 3
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n
out = {}
us = timeit(lambda: TF.flow_loss_grad(mov, tgt, flow, 0.5, 0.5)); out["flow_node_mse+ncc"] = {"us": us, "alg_GBps": 52 * vox / us / 1e3}
us = timeit(lambda: TF.flow_loss_grad(mov, tgt, flow, 1.0, 0.0)); out["flow_node_mse"] = {"us": us, "alg_GBps": 52 * vox / us / 1e3}
us = timeit(lambda: TF.warp_flow(mov, flow)); out["warp_flow"] = {"us": us, "alg_GBps": 20 * vox / us / 1e3}
th = torch.tensor([[1.01, .02, -.01, .01], [-.02, .99, .01, 0.], [.01, -.01, 1., .02]], device=dev)
us = timeit(lambda: TF.warp_affine(th, mov)); out["warp_affine"] = {"us": us, "alg_GBps": 8 * vox / us / 1e3}
for opt, bpv in (("sgd", 52), ("adam", 100)):
    prob = TF.DirectFlowProblem(mov, tgt, 100000, optimiser=opt)
    us = timeit(lambda: prob.run(1, 0.05, 0.5, 0.5, 2.0)); out["direct_flow_epoch_" + opt] = {"us": us, "alg_GBps": bpv * vox / us / 1e3, "bytes_per_voxel": bpv}
print(json.dumps({"size": S, "results": out}))
