"""One launch of every kernel of libaisp_b200.so at the bench sizes, for `ncu --set full` (profiles/README.md):

    ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_all_r02 \
        python scripts/profile_all_kernels.py

Sizes: 64 x 3 x 512 x 512 (BASELINE configs[1]) for everything, plus USM / per-pixel at 8 x 3 x 2160 x 3840
(configs[3]).  Each section prints its name so that the report's launch order can be read back."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adaptiveisp_b200 import _lib, functional as AF  # noqa: E402
from adaptiveisp_b200.synthetic import lod_batch  # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()
st = torch.cuda.current_stream(dev).cuda_stream
ck = _lib.check


def params_for(op, B):
    P = torch.zeros((B, 24), device=dev)
    vals = {AF.OP_EXPOSURE: [0.6], AF.OP_GAMMA: [0.7], AF.OP_CCM: [1.6, -0.4, -0.2, -0.3, 1.5, -0.2, -0.1, -0.5, 1.6],
            AF.OP_SHARPEN: [3.0], AF.OP_NLM: [0.3], AF.OP_TONE: [0.6, 0.8, 1.0, 1.2, 1.4, 1.6, 1.8, 1.9], AF.OP_CONTRAST: [0.4],
            AF.OP_SATPLUS: [0.5], AF.OP_WNB: [0.4], AF.OP_WB: [1.1, 0.9, 1.2], AF.OP_USM: [1.0, 1.2],
            AF.OP_COLOR: [0.9 + 0.2 * k / 23 for k in range(24)], AF.OP_SHARPEN_V2: [1.5]}[op]
    P[:, :len(vals)] = torch.tensor(vals, device=dev)
    return P


def section(name):
    torch.cuda.synchronize()
    print("section:", name, flush=True)


def run(B, H, W, tag, ops_list):
    img = lod_batch(B, H, W, seed=7, device=dev, letterbox=(H == 512))
    g = torch.randn_like(img)
    out, gi, gy, stash = torch.empty_like(img), torch.empty_like(img), torch.empty_like(img), torch.empty_like(img)
    wsum = torch.empty((B, 1, H, W), device=dev)
    gP = torch.zeros((B, 24), device=dev)
    sc = _lib.scratch(B, H, W, dev)
    for op in ops_list:
        P = params_for(op, B)
        od = torch.full((B,), op, dtype=torch.int32, device=dev)
        fam = AF.family_of(op)
        section(f"{tag} op {op} {fam} fwd / bwd(params) / bwd(params+img)")
        if fam == AF.FAMILY_POINTWISE:
            ck(L.aisp_pointwise_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), od.data_ptr(), None, B, H, W, 1, 1, st), "f")
            ck(L.aisp_pointwise_bwd(img.data_ptr(), g.data_ptr(), P.data_ptr(), od.data_ptr(), B, H, W, 1, gP.data_ptr(), None,
                                    sc.data_ptr(), sc.numel(), st), "b")
            ck(L.aisp_pointwise_bwd(img.data_ptr(), g.data_ptr(), P.data_ptr(), od.data_ptr(), B, H, W, 1, gP.data_ptr(),
                                    gi.data_ptr(), sc.data_ptr(), sc.numel(), st), "bi")
        elif fam == AF.FAMILY_SHARPEN:
            ck(L.aisp_sharpen_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), od.data_ptr(), B, H, W, st), "f")
            ck(L.aisp_sharpen_bwd(img.data_ptr(), g.data_ptr(), P.data_ptr(), od.data_ptr(), B, H, W, gP.data_ptr(), None, None,
                                  sc.data_ptr(), sc.numel(), st), "b")
            ck(L.aisp_sharpen_bwd(img.data_ptr(), g.data_ptr(), P.data_ptr(), od.data_ptr(), B, H, W, gP.data_ptr(),
                                  gi.data_ptr(), gy.data_ptr(), sc.data_ptr(), sc.numel(), st), "bi")
        else:
            ck(L.aisp_nlm_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), od.data_ptr(), B, H, W, stash.data_ptr(),
                              wsum.data_ptr(), st), "f")
            ck(L.aisp_nlm_bwd(g.data_ptr(), stash.data_ptr(), od.data_ptr(), B, H, W, gP.data_ptr(), sc.data_ptr(), sc.numel(), st), "b")
    return img, g, out, gi, sc


# ---- single filters at the bench size (every family, every per-pixel op) and the 4K stencils
img, g, out, gi, sc = run(64, 512, 512, "64x512x512", list(range(13)))
B, H, W = 64, 512, 512

section("NLM image gradient (rare path) on 4 frames")
P = params_for(AF.OP_NLM, 4)
od = torch.full((4,), AF.OP_NLM, dtype=torch.int32, device=dev)
ws = torch.empty((4, 1, H, W), device=dev)
ck(L.aisp_nlm_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), od.data_ptr(), 4, H, W, None, ws.data_ptr(), st), "f")
ck(L.aisp_nlm_bwd_img(img.data_ptr(), out.data_ptr(), ws.data_ptr(), g.data_ptr(), P.data_ptr(), od.data_ptr(), 4, H, W,
                      gi.data_ptr(), st), "bi")

section("fused sequence E->G->WB->CCM: forward, backward(params), backward(params+img)")
seq = [AF.OP_EXPOSURE, AF.OP_GAMMA, AF.OP_WB, AF.OP_CCM]
Pc = torch.stack([params_for(o, B) for o in seq], 1).contiguous()
Pc[:, 0, 0], Pc[:, 1, 0] = 0.09012079, 0.38566995
oc = torch.tensor([seq] * B, dtype=torch.int32, device=dev)
gPc = torch.zeros_like(Pc)
ck(L.aisp_pointwise_fwd(img.data_ptr(), out.data_ptr(), Pc.data_ptr(), oc.data_ptr(), None, B, H, W, 4, 1, st), "cf")
ck(L.aisp_pointwise_chain_bwd(img.data_ptr(), g.data_ptr(), Pc.data_ptr(), oc.data_ptr(), None, B, H, W, 4, 1, gPc.data_ptr(), None,
                              sc.data_ptr(), sc.numel(), st), "cb")
ck(L.aisp_pointwise_chain_bwd(img.data_ptr(), g.data_ptr(), Pc.data_ptr(), oc.data_ptr(), None, B, H, W, 4, 1, gPc.data_ptr(),
                              gi.data_ptr(), sc.data_ptr(), sc.numel(), st), "cbi")

section("sequence launch set: E->G->WB->CCM->Shr in one launch; BW->NLM->T; with 64x64 block means")
for seq in ([AF.OP_EXPOSURE, AF.OP_GAMMA, AF.OP_WB, AF.OP_CCM, AF.OP_SHARPEN], [AF.OP_WNB, AF.OP_NLM, AF.OP_TONE]):
    Ps = torch.stack([params_for(o, B) for o in seq], 1).contiguous()
    os_ = torch.tensor([seq] * B, dtype=torch.int32, device=dev)
    AF.sequence_forward(img, Ps, os_, None, True, down_hw=(64, 64))

section("select-apply (agent semantics): heterogeneous batch, forward with twin-less block means, backward")
ops_h = torch.tensor([i % 10 for i in range(B)], dtype=torch.int32, device=dev)
Ph = torch.stack([params_for(int(o), 1)[0] for o in ops_h.tolist()], 0).contiguous().requires_grad_(True)
y, _, down = AF.apply_ops(img, Ph, ops_h, clip=True, down_hw=(64, 64))
(y * g).sum().backward()

section("block mean 512->64, value statistics, device-side selection, regressors")
d64 = AF.block_mean(img, (64, 64))
AF.value_stats(d64)
F = 10
pdf = torch.softmax(torch.randn((B, F), device=dev), dim=1)
pk = torch.randn((B, F, 24), device=dev, requires_grad=True)
rows = AF.select_rows(pdf, torch.rand((B, 1), device=dev), torch.zeros((B, 3 + F), device=dev), pk,
                      torch.arange(F, dtype=torch.int32, device=dev), AF.SELECT_SAMPLE)[0]
rows.sum().backward()
from adaptiveisp_b200 import filters as Fm  # noqa: E402
from adaptiveisp_b200.config import make_cfg  # noqa: E402
cfg = make_cfg()
mods = [c(cfg, predict=True).to(dev) for c in cfg.filters]
pred = Fm.BankPredictor(mods)
feats = torch.randn((B, cfg.feature_extractor_dims), device=dev) * 0.05
pred(feats).sum().backward()

section("filter bank, 10 cfg.filters: forward + backward (the bench step)")
bank_ops = (ctypes.c_int32 * 10)(*[m.OP for m in mods])
P_all = torch.stack([params_for(m.OP, B) for m in mods], 1).contiguous()
out_all = torch.empty((B, 10, 3, H, W), device=dev)
gout_all = torch.randn_like(out_all)
gP_all = torch.zeros((B, 10, 24), device=dev)
stash = torch.empty_like(img)
sc_all = _lib.scratch(B * 10, H, W, dev)
ck(L.aisp_bank_fwd(img.data_ptr(), out_all.data_ptr(), P_all.data_ptr(), bank_ops, B, 10, H, W, 1, stash.data_ptr(), st), "bf")
ck(L.aisp_bank_bwd(img.data_ptr(), gout_all.data_ptr(), P_all.data_ptr(), bank_ops, B, 10, H, W, 1, stash.data_ptr(), gP_all.data_ptr(),
                   sc_all.data_ptr(), sc_all.numel(), st), "bb")
del out_all, gout_all, img, g, out, gi

section("4K (8 x 2160 x 3840): BW, USM, sharpen")
run(8, 2160, 3840, "8x2160x3840", [AF.OP_WNB, AF.OP_USM, AF.OP_SHARPEN])
torch.cuda.synchronize()
print("profile_all_kernels: done")
