"""Does the DRAM-bound part of the bank forward overlap with the issue-bound NLM when they run on two streams?"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from adaptiveisp_b200 import _lib, functional as AF
from adaptiveisp_b200.config import make_cfg
from adaptiveisp_b200.synthetic import lod_batch

dev = torch.device("cuda:0")
L = _lib.lib()
B, H, W = 64, 512, 512
cfg = make_cfg()
flts = [c(cfg, predict=True).to(dev) for c in cfg.filters]
img = lod_batch(B, H, W, seed=1235, device=dev)
feats = torch.randn((B, cfg.feature_extractor_dims), device=dev) * 0.05
with torch.no_grad():
    packed = [AF.pack_params(f.filter_param_regressor(f.extract_parameters(feats)[0]), f.get_num_filter_parameters()) for f in flts]
light = [i for i, f in enumerate(flts) if f.OP != AF.OP_NLM]
heavy = [i for i, f in enumerate(flts) if f.OP == AF.OP_NLM]
Pl = torch.stack([packed[i] for i in light], 1).contiguous()
Ph = torch.stack([packed[i] for i in heavy], 1).contiguous()
ol = (ctypes.c_int32 * len(light))(*[flts[i].OP for i in light])
oh = (ctypes.c_int32 * 1)(AF.OP_NLM)
out_l = torch.empty((B, len(light), 3, H, W), device=dev)
out_h = torch.empty((B, 1, 3, H, W), device=dev)
stash = torch.empty_like(img)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(concurrent, nlm_first):
    cur = torch.cuda.current_stream()
    a = s1 if concurrent else cur
    b = s2 if concurrent else cur
    if concurrent:
        s1.wait_stream(cur); s2.wait_stream(cur)
    def light_call():
        L.aisp_bank_fwd(img.data_ptr(), out_l.data_ptr(), Pl.data_ptr(), ol, B, len(light), H, W, 1, None, a.cuda_stream)
    def heavy_call():
        L.aisp_bank_fwd(img.data_ptr(), out_h.data_ptr(), Ph.data_ptr(), oh, B, 1, H, W, 1, stash.data_ptr(), b.cuda_stream)
    if nlm_first:
        heavy_call(); light_call()
    else:
        light_call(); heavy_call()
    if concurrent:
        cur.wait_stream(s1); cur.wait_stream(s2)


for conc in (False, True):
    for nf in (False, True):
        for _ in range(3):
            run(conc, nf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run(conc, nf)
        e1.record()
        torch.cuda.synchronize()
        print(f"concurrent={conc} nlm_first={nf}: {e0.elapsed_time(e1) / 10:.3f} ms")
