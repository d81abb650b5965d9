import sys, torch, numpy as np
sys.path.insert(0, '.')
from oracle import isp_oracle as O
from tests import cases
from adaptiveisp_b200 import filters as F
from adaptiveisp_b200.config import make_cfg
cfg = make_cfg(); dev = torch.device('cuda:0')
torch.set_num_threads(8)
flt = F.DenoiseFilter(cfg).to(dev)
for (B,H,W,seed) in [(1,67,341,11),(1,64,64,1),(1,40,56,2),(2,33,50,11)]:
    img = cases.edge_image(B,H,W,seed=seed,in_range=True)
    _, param = cases.params_for(O.OP_NLM,B,seed=seed)
    g = cases.grad_out(img.shape, seed=seed)
    pc = param.clone().requires_grad_(True)
    yc = O.forward(O.OP_NLM, img, pc); (yc*g).sum().backward()
    # fp64 reference
    p64 = param.double().clone().requires_grad_(True)
    y64 = O.nlm_gray(img.double(), p64)
    (y64*g.double()).sum().backward()
    pd = param.to(dev).requires_grad_(True)
    yd = flt.forward(img.to(dev), specified_parameter=pd)[0]; (yd*g.to(dev)).sum().backward()
    print((B,H,W), 'h', param.flatten().tolist(), 'out err vs f32', float((yd.cpu()-yc).abs().max()), 'vs f64', float((yd.cpu().double()-y64).abs().max()),
          'ref32 vs f64', float((yc.double()-y64).abs().max()))
    print('   grad cuda', pd.grad.flatten().tolist(), 'ref32', pc.grad.flatten().tolist(), 'ref64', p64.grad.flatten().tolist())
