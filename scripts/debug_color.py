import sys, torch, numpy as np
sys.path.insert(0, '.')
from oracle import isp_oracle as O
from tests import cases
from adaptiveisp_b200 import filters as F
from adaptiveisp_b200.config import make_cfg
cfg = make_cfg(); dev = torch.device('cuda:0')
for op in (O.OP_COLOR, O.OP_TONE):
  for (B,H,W) in [(2,96,136),(1,64,64),(1,64,65),(1,32,32),(2,20,24)]:
    flt = {O.OP_COLOR: F.ColorFilter, O.OP_TONE: F.ToneFilter}[op](cfg).to(dev)
    img = cases.edge_image(B,H,W,seed=11)
    _, param = cases.params_for(op,B,seed=11)
    g = cases.grad_out(img.shape, seed=11)
    xc = img.clone().requires_grad_(True); pc = param.clone().requires_grad_(True)
    (O.forward(op, xc, pc)*g).sum().backward()
    xd = img.to(dev).requires_grad_(True); pd = param.to(dev).requires_grad_(True)
    (flt.forward(xd, specified_parameter=pd)[0]*g.to(dev)).sum().backward()
    d = (xd.grad.cpu()-xc.grad).abs()
    bad = (d > 1e-4*xc.grad.abs().max()).nonzero()
    print(O.OP_NAMES[op], (B,H,W), 'max', float(d.max()), 'nbad', len(bad), bad[:6].tolist())
    for idx in bad[:4].tolist():
        b,c,y,x = idx
        print('   x=', float(img[b,c,y,x]), 'ref', float(xc.grad[b,c,y,x]), 'got', float(xd.grad[b,c,y,x].cpu()), 'g', float(g[b,c,y,x]))
