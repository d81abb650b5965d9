"""Where does an Agent training step spend its time on a B200?  (caller-level view)"""
import sys, time, torch
sys.path.insert(0, ".")
from adaptiveisp_b200.agent import Agent
from adaptiveisp_b200.config import make_cfg
from adaptiveisp_b200.synthetic import lod_batch
dev = torch.device("cuda:0")
cfg = make_cfg()
for B in (64, 8):
    agent = Agent(cfg, shape=(16, 64, 64), device=dev).to(dev)
    agent.train()
    x = lod_batch(B, 512, 512, seed=1, device=dev)
    z = torch.rand((B, cfg.z_dim), device=dev)
    states = torch.zeros((B, cfg.num_state_dim), device=dev)
    g = torch.randn_like(x)
    def step():
        (xo, ns, sur, pen), dbg, _ = agent((x, z, states), 0.5)
        loss = (xo * g).sum() * 1e-6 + sur.sum() + pen.sum()
        loss.backward()
        return xo
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10): step()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"B={B}: eager agent fwd+bwd: GPU {e0.elapsed_time(e1)/10:.2f} ms/step, CPU launch {(t1-t0)*100:.2f} ms/step")
    # CUDA graph of the whole step
    agent.zero_grad(set_to_none=True)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): step()
    torch.cuda.current_stream().wait_stream(s)
    agent.zero_grad(set_to_none=True)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = step()
    torch.cuda.synchronize()
    for _ in range(3): gr.replay()
    e0.record()
    for _ in range(10): gr.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"B={B}: graphed agent fwd+bwd: GPU {e0.elapsed_time(e1)/10:.2f} ms/step")
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        gr.replay(); torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:8]
    for r in rows:
        print(f"    {r.key[:70]:70s} {r.device_time_total/1e3:8.3f} ms  x{r.count}")
