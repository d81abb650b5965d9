"""Filter bank vs per-filter launches: all 10 cfg.filters fwd+bwd on 64x3x512x512 (device-resident)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptiveisp_b200 import _lib, functional as AF
from adaptiveisp_b200.config import make_cfg
from adaptiveisp_b200.synthetic import lod_batch

dev = torch.device("cuda:0")
L = _lib.lib()
B, H, W = 64, 512, 512
cfg = make_cfg()
flts = [c(cfg, predict=True).to(dev) for c in cfg.filters]
Fn = len(flts)
img = lod_batch(B, H, W, seed=1235, device=dev)
feats = torch.randn((B, cfg.feature_extractor_dims), device=dev) * 0.05
with torch.no_grad():
    packed = [AF.pack_params(f.filter_param_regressor(f.extract_parameters(feats)[0]), f.get_num_filter_parameters())
              for f in flts]
P = torch.stack(packed, 1).contiguous()
import ctypes
ops = (ctypes.c_int32 * Fn)(*[f.OP for f in flts])
out = torch.empty((B, Fn, 3, H, W), device=dev)
gout = torch.randn((B, Fn, 3, H, W), device=dev)
stash = torch.empty_like(img)
gP = torch.zeros_like(P)
sc = _lib.scratch(B * Fn, H, W, dev)
st = torch.cuda.current_stream(dev).cuda_stream


def timed(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def bank_fwd():
    _lib.check(L.aisp_bank_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), ops, B, Fn, H, W, 1,
                               stash.data_ptr(), st), "bank fwd")


def bank_bwd():
    _lib.check(L.aisp_bank_bwd(img.data_ptr(), gout.data_ptr(), P.data_ptr(), ops, B, Fn, H, W, 1,
                               stash.data_ptr(), gP.data_ptr(), sc.data_ptr(), sc.numel(), st), "bank bwd")


tf, tb = timed(bank_fwd), timed(bank_bwd)
print(f"bank fwd {tf:.3f} ms  bwd {tb:.3f} ms  total {tf + tb:.3f} ms  -> {Fn * B * H * W / 1e6 / ((tf + tb) / 1e3):.0f} MP/s")
# family by family: a bank holding only that family's filters
for name in ("pointwise", "sharpen", "nlm"):
    keep = [i for i, f in enumerate(flts) if AF.family_of(f.OP) == name]
    o = (ctypes.c_int32 * len(keep))(*[flts[i].OP for i in keep])
    Pk = P[:, keep].contiguous()
    nk = len(keep)
    f1 = timed(lambda: L.aisp_bank_fwd(img.data_ptr(), out.data_ptr(), Pk.data_ptr(), o, B, nk, H, W, 1,
                                       stash.data_ptr(), st))
    b1 = timed(lambda: L.aisp_bank_bwd(img.data_ptr(), gout.data_ptr(), Pk.data_ptr(), o, B, nk, H, W, 1,
                                       stash.data_ptr(), gP.data_ptr(), sc.data_ptr(), sc.numel(), st))
    gb = nk * B * H * W * 24 / 1e9
    print(f"  only {name:9s} ({nk}): fwd {f1:.3f} ms ({gb / f1 * 1e3:.0f} GB/s algorithmic)  bwd {b1:.3f} ms ({gb / b1 * 1e3:.0f} GB/s)")

# per-filter loop (the round-1 bench path)
o1 = [torch.full((B,), f.OP, dtype=torch.int32, device=dev) for f in flts]
out1 = torch.empty_like(img)
g1 = gout[:, 0].contiguous()
gP1 = torch.zeros((B, 24), device=dev)


def loop():
    for i, f in enumerate(flts):
        fam = AF.family_of(f.OP)
        Pi = packed[i]
        if fam == "pointwise":
            L.aisp_pointwise_fwd(img.data_ptr(), out1.data_ptr(), Pi.data_ptr(), o1[i].data_ptr(), None, B, H, W, 1, 1, st)
            L.aisp_pointwise_bwd(img.data_ptr(), g1.data_ptr(), Pi.data_ptr(), o1[i].data_ptr(), B, H, W, 1,
                                 gP1.data_ptr(), None, sc.data_ptr(), sc.numel(), st)
        elif fam == "sharpen":
            L.aisp_sharpen_fwd(img.data_ptr(), out1.data_ptr(), Pi.data_ptr(), o1[i].data_ptr(), B, H, W, st)
            L.aisp_sharpen_bwd(img.data_ptr(), g1.data_ptr(), Pi.data_ptr(), o1[i].data_ptr(), B, H, W,
                               gP1.data_ptr(), None, None, sc.data_ptr(), sc.numel(), st)
        else:
            L.aisp_nlm_fwd(img.data_ptr(), out1.data_ptr(), Pi.data_ptr(), o1[i].data_ptr(), B, H, W, stash.data_ptr(), None, st)
            L.aisp_nlm_bwd(g1.data_ptr(), stash.data_ptr(), o1[i].data_ptr(), B, H, W, gP1.data_ptr(), sc.data_ptr(), sc.numel(), st)


tl = timed(loop)
print(f"per-filter loop {tl:.3f} ms -> {Fn * B * H * W / 1e6 / (tl / 1e3):.0f} MP/s")
