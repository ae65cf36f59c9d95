"""torch.profiler table of one Register(mode='flow') U-Net epoch at S^3."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn as nn
import torchregister_b200 as tr
from torchregister_b200.synth import make_pair
from torch.profiler import profile, ProfilerActivity
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda:0"
mov, tgt = make_pair((S, S, S), "flow", device=dev)
torch.manual_seed(0)
fr = tr.flow_register((S, S, S), mode="bilinear", n=32, lr=1e-3, max_epochs=1, criterions=[nn.MSELoss(), tr.NCCLoss()],
                      weights=[0.5, 0.5], stop_crit=-1.0).to(dev)
fr.optimize(mov, tgt, dev, debug=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    fr.optimize(mov, tgt, dev, debug=False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key in ("aten::convolution_backward", "aten::cudnn_convolution", "aten::cudnn_convolution_transpose", "aten::conv3d")]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:40]:
    print("%-36s %8.2f ms x%d  %s" % (e.key, e.device_time_total / 1e3, e.count, [s_ for s_ in e.input_shapes if s_][:3]))
