import sys, os, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchregister_b200 as tr
g = dict(np.load("tests/golden/flowreg2d.npz"))
sd = {k[4+6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd::")}
mov = torch.from_numpy(g["moving"])
torch.backends.cudnn.allow_tf32 = False
nc = tr.Attention_UNet((160,168), "bilinear", in_c=1, n=32); nc.load_state_dict(sd)
ng = tr.Attention_UNet((160,168), "bilinear", in_c=1, n=32); ng.load_state_dict(sd); ng = ng.cuda()
def trace(net, x, dev):
    out = {}
    y1 = net.layer1(x); out['y1']=y1
    y2 = net.layer2(net.maxpool(y1)); out['y2']=y2
    y3 = net.layer3(net.maxpool(y2)); out['y3']=y3
    y4 = net.layer4(net.maxpool(y3)); out['y4']=y4
    y = net.layer5(net.maxpool(y4)); out['y5']=y
    a = net.skip4.input_filter(y4); b = net.skip4.gate_filter(y); out['a']=a; out['b']=b
    from torchregister_b200.utils import padNd
    if a.shape[-1] < b.shape[-1]: a = padNd(a,b,dev)
    elif a.shape[-1] > b.shape[-1]: b = padNd(b,a,dev)
    w = torch.sigmoid(net.skip4.psi(F.relu(a+b))); out['w']=w
    wi = F.interpolate(w, size=y4.shape[2:], mode='nearest'); out['wi']=wi
    g4 = net.skip4.bnorm(y4*wi); out['g4']=g4
    return out
with torch.no_grad():
    tc = trace(nc, mov, 'cpu'); tg = trace(ng, mov.cuda(), 'cuda')
for k in tc:
    print(k, tuple(tc[k].shape), "max diff %.3e  scale %.3e" % ((tg[k].cpu()-tc[k]).abs().max().item(), tc[k].abs().max().item()))
