"""Register(mode='flow') U-Net epoch at 256^3 under different cuDNN settings (all reference-compatible: same fp32 model)."""
import sys, os, json, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn as nn
import torchregister_b200 as tr
from torchregister_b200.synth import make_pair
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda:0"
mov, tgt = make_pair((S, S, S), "flow", device=dev)
def run(tag, **flags):
    torch.manual_seed(0)
    with torch.backends.cudnn.flags(**flags):
        fr = tr.flow_register((S, S, S), mode="bilinear", n=32, lr=1e-3, max_epochs=2, criterions=[nn.MSELoss(), tr.NCCLoss()],
                              weights=[0.5, 0.5], stop_crit=-1.0).to(dev)
        fr.optimize(mov, tgt, dev, debug=False)
        torch.cuda.synchronize()
        fr.max_epochs = 3
        t0 = time.perf_counter()
        fr.optimize(mov, tgt, dev, debug=False)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
    print("%-28s %.1f ms/epoch  losses %s" % (tag, ms, ["%.5f" % v for v in fr.losses[:3]]), flush=True)
    del fr; torch.cuda.empty_cache()
run("default", enabled=True, benchmark=False, deterministic=False, allow_tf32=True)
run("benchmark", enabled=True, benchmark=True, deterministic=False, allow_tf32=True)
run("benchmark, no tf32", enabled=True, benchmark=True, deterministic=False, allow_tf32=False)
run("cudnn off", enabled=False, benchmark=False, deterministic=False, allow_tf32=True)
