"""Timing of the HBM-bound kernels VERDICT r1 lists below 0.70 of the HBM peak, through the C ABI:
3x3 sharpen / SharpenV2 / USM backward (64 x 512^2 and 8 x 2160 x 3840), Saturation+ and Tone forward / backward.
    python scripts/micro/laggard_bench.py [--iters 20] [--only usm4k]
One JSON line per case: ms, GB/s on the algorithmic bytes (24 B/px per pass: 12 read + 12 written forward,
image + upstream gradient read backward), fraction of the measured HBM peak."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from adaptiveisp_b200 import _lib, functional as AF  # noqa: E402
from adaptiveisp_b200.synthetic import lod_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    L = _lib.lib()
    peak = 6551.7
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    st = torch.cuda.current_stream(dev).cuda_stream
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(args.iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / args.iters

    vals = {AF.OP_SHARPEN: [3.0], AF.OP_SHARPEN_V2: [1.5], AF.OP_USM: [1.0, 1.2], AF.OP_SATPLUS: [0.5],
            AF.OP_TONE: [0.6, 0.8, 1.0, 1.2, 1.4, 1.6, 1.8, 1.9], AF.OP_EXPOSURE: [0.6]}
    for tag, (B, H, W) in (("512", (64, 512, 512)), ("4k", (8, 2160, 3840))):
        img = lod_batch(B, H, W, seed=7, device=dev, letterbox=(H == 512))
        g = torch.randn_like(img)
        out, gi, gy = torch.empty_like(img), torch.empty_like(img), torch.empty_like(img)
        gP = torch.zeros((B, 24), device=dev)
        sc = _lib.scratch(B, H, W, dev)
        nbytes = 24.0 * B * H * W
        for name, op in (("shr", AF.OP_SHARPEN), ("shr2", AF.OP_SHARPEN_V2), ("usm", AF.OP_USM), ("satp", AF.OP_SATPLUS),
                         ("tone", AF.OP_TONE), ("exp", AF.OP_EXPOSURE)):
            case = name + tag
            if args.only and args.only not in case:
                continue
            P = torch.zeros((B, 24), device=dev)
            P[:, :len(vals[op])] = torch.tensor(vals[op], device=dev)
            od = torch.full((B,), op, dtype=torch.int32, device=dev)
            if AF.family_of(op) == AF.FAMILY_SHARPEN:
                f = lambda: _lib.check(L.aisp_sharpen_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), od.data_ptr(), B, H, W, st), "f")
                b = lambda: _lib.check(L.aisp_sharpen_bwd(img.data_ptr(), g.data_ptr(), P.data_ptr(), od.data_ptr(), B, H, W,
                                                          gP.data_ptr(), None, None, sc.data_ptr(), sc.numel(), st), "b")
                bi = lambda: _lib.check(L.aisp_sharpen_bwd(img.data_ptr(), g.data_ptr(), P.data_ptr(), od.data_ptr(), B, H, W,
                                                           gP.data_ptr(), gi.data_ptr(), gy.data_ptr(), sc.data_ptr(), sc.numel(), st), "bi")
            else:
                f = lambda: _lib.check(L.aisp_pointwise_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), od.data_ptr(), None, B, H, W,
                                                            1, 1, st), "f")
                b = lambda: _lib.check(L.aisp_pointwise_bwd(img.data_ptr(), g.data_ptr(), P.data_ptr(), od.data_ptr(), B, H, W, 1,
                                                            gP.data_ptr(), None, sc.data_ptr(), sc.numel(), st), "b")
                bi = lambda: _lib.check(L.aisp_pointwise_bwd(img.data_ptr(), g.data_ptr(), P.data_ptr(), od.data_ptr(), B, H, W, 1,
                                                             gP.data_ptr(), gi.data_ptr(), sc.data_ptr(), sc.numel(), st), "bi")
            tf, tb, tbi = timed(f), timed(b), timed(bi)
            print(json.dumps({"case": case, "fwd_ms": round(tf, 4), "bwd_ms": round(tb, 4), "bwd_gimg_ms": round(tbi, 4),
                              "fwd_frac": round(nbytes / tf / 1e6 / peak, 3), "bwd_frac": round(nbytes / tb / 1e6 / peak, 3),
                              "gP_checksum": float(gP.double().abs().sum())}), flush=True)
        del img, g, out, gi, gy, sc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
