import sys, time, torch
sys.path.insert(0, '.')
from adaptiveisp_b200.config import make_cfg
from adaptiveisp_b200.synthetic import lod_batch
dev = torch.device('cuda:0'); cfg = make_cfg()
B, H, W = 64, 512, 512
img = lod_batch(B, H, W, seed=1, device=dev)
gout = torch.randn_like(img)
flts = [c(cfg, predict=True).to(dev) for c in cfg.filters]
ft = torch.randn((B, 4096), device=dev) * 0.05
def isp_step(x, ft):
    last = None
    for f in flts:
        y, _, _ = f(x, ft); y.backward(gout); last = y
    return last
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
# CUDA graph of the whole fwd+bwd step
for f in flts: f.zero_grad(set_to_none=True)
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): isp_step(img, ft)
torch.cuda.current_stream().wait_stream(s)
for f in flts: f.zero_grad(set_to_none=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = isp_step(img, ft)
torch.cuda.synchronize()
for _ in range(3): g.replay()
torch.cuda.synchronize()
t0 = time.perf_counter(); e0.record()
for _ in range(20): g.replay()
e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
print("graphed: CPU launch %.3f ms/step, GPU %.2f ms/step" % ((t1 - t0) * 50, e0.elapsed_time(e1) / 20))

for _ in range(3): isp_step(img, ft)
torch.cuda.synchronize()
t0 = time.perf_counter(); e0.record()
for _ in range(20): isp_step(img, ft)
e1.record(); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print("eager class API: CPU launch %.2f ms/step, GPU %.2f ms/step, wall %.2f ms/step" % ((t1 - t0) * 50, e0.elapsed_time(e1) / 20, (t2 - t0) * 50))
