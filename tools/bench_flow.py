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
for opt, bpv2, bpv1 in (("sgd", 52, 32), ("adam", 100, 80)):
    prob = TF.DirectFlowProblem(mov, tgt, 1000000, optimiser=opt)
    us = timeit(lambda: prob.run_two_pass(1, 0.05, 0.5, 0.5, 2.0)); out["direct_flow_two_pass_" + opt] = {"us": us, "alg_GBps": bpv2 * vox / us / 1e3, "bytes_per_voxel": bpv2}
    for name, w, sm in (("mse+ncc_smooth", (0.5, 0.5), 2.0), ("mse_smooth", (1.0, 0.0), 2.0), ("mse+ncc", (0.5, 0.5), 0.0), ("mse", (1.0, 0.0), 0.0)):
        prob = TF.DirectFlowProblem(mov, tgt, 1000000, optimiser=opt)
        us = timeit(lambda: prob.run(10, 0.05, w[0], w[1], sm)) / 10
        out["direct_flow_fused_%s_%s" % (opt, name)] = {"us": us, "alg_GBps": bpv1 * vox / us / 1e3, "bytes_per_voxel": bpv1}
lib = TF._lib.load()
lib.trb_flow_direct_set_path(1)
for opt, bpv1 in (("sgd", 32), ("adam", 80)):
    for name, w, sm in (("mse+ncc_smooth", (0.5, 0.5), 2.0), ("mse_smooth", (1.0, 0.0), 2.0)):
        prob = TF.DirectFlowProblem(mov, tgt, 1000000, optimiser=opt)
        us = timeit(lambda: prob.run(10, 0.05, w[0], w[1], sm)) / 10
        out["direct_flow_fused_noTMA_%s_%s" % (opt, name)] = {"us": us, "alg_GBps": bpv1 * vox / us / 1e3, "bytes_per_voxel": bpv1}
lib.trb_flow_direct_set_path(0)
print(json.dumps({"size": S, "results": out}))
