"""Timing of the fused sequence kernels (forward / backward) at the bench size, through the C ABI.
    python scripts/micro/chain_bench.py [--iters 20]
Prints one JSON line per case: ms, GB/s on the algorithmic bytes (24 B/px each way), fraction of the
measured HBM peak.  Cases: E->G->WB->CCM (isp/filters.py:753-815), a 5-stage runtime-penalty style
sequence, a single stage, with and without the image gradient."""
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
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--cases", type=int, default=0, help="only the first N cases (0: all)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    L = _lib.lib()
    B, H, W = args.batch, 512, 512
    peak = 6551.7
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    img = lod_batch(B, H, W, seed=1240, device=dev)
    g = torch.randn_like(img)
    o = torch.empty_like(img)
    gi = torch.empty_like(img)
    st = torch.cuda.current_stream(dev).cuda_stream
    sc = _lib.scratch(B, H, W, dev)
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

    def params(seq):
        S = len(seq)
        P = torch.zeros((B, S, 24), device=dev)
        for k, op in enumerate(seq):
            if op == AF.OP_EXPOSURE:
                P[:, k, 0] = 0.09012079
            elif op == AF.OP_GAMMA:
                P[:, k, 0] = 0.38566995
            elif op == AF.OP_WB:
                P[:, k, :3] = torch.tensor([2.4052505, 1.2233436, 1.8800205], device=dev)
            elif op == AF.OP_CCM:
                P[:, k, :9] = torch.tensor([1.6, -0.4, -0.2, -0.3, 1.5, -0.2, -0.1, -0.5, 1.6], device=dev)
            elif op in (AF.OP_TONE,):
                P[:, k, :8] = torch.linspace(0.6, 1.8, 8, device=dev)
            elif op == AF.OP_COLOR:
                P[:, k, :24] = torch.linspace(0.9, 1.1, 24, device=dev)
            else:
                P[:, k, 0] = 0.4
        return P

    cases = [
        ("E_G_WB_CCM", [AF.OP_EXPOSURE, AF.OP_GAMMA, AF.OP_WB, AF.OP_CCM]),
        ("E", [AF.OP_EXPOSURE]),
        ("CCM", [AF.OP_CCM]),
        ("E_G", [AF.OP_EXPOSURE, AF.OP_GAMMA]),
        ("WB_CCM_T_Ct_BW", [AF.OP_WB, AF.OP_CCM, AF.OP_TONE, AF.OP_CONTRAST, AF.OP_WNB]),
        ("G_T_S+_Ct_E_CCM", [AF.OP_GAMMA, AF.OP_TONE, AF.OP_SATPLUS, AF.OP_CONTRAST, AF.OP_EXPOSURE, AF.OP_CCM]),
        ("T_C", [AF.OP_TONE, AF.OP_COLOR]),
    ]
    npx = B * H * W
    if args.cases:
        cases = cases[:args.cases]
    for name, seq in cases:
        S = len(seq)
        P = params(seq)
        gP = torch.zeros_like(P)
        ops = torch.tensor([seq] * B, dtype=torch.int32, device=dev)
        t_f = timed(lambda: _lib.check(L.aisp_pointwise_fwd(img.data_ptr(), o.data_ptr(), P.data_ptr(), ops.data_ptr(), None,
                                                           B, H, W, S, 1, st), "fwd"))
        t_b = timed(lambda: _lib.check(L.aisp_pointwise_chain_bwd(img.data_ptr(), g.data_ptr(), P.data_ptr(), ops.data_ptr(),
                                                                 None, B, H, W, S, 1, gP.data_ptr(), None, sc.data_ptr(),
                                                                 sc.numel(), st), "bwd"))
        t_bi = timed(lambda: _lib.check(L.aisp_pointwise_chain_bwd(img.data_ptr(), g.data_ptr(), P.data_ptr(), ops.data_ptr(),
                                                                  None, B, H, W, S, 1, gP.data_ptr(), gi.data_ptr(),
                                                                  sc.data_ptr(), sc.numel(), st), "bwd+gimg"))
        row = {"case": name, "S": S, "fwd_ms": round(t_f, 4), "bwd_ms": round(t_b, 4), "bwd_gimg_ms": round(t_bi, 4),
               "fwd_frac": round(24 * npx / 1e6 / t_f / peak, 3), "bwd_frac": round(24 * npx / 1e6 / t_b / peak, 3),
               "bwd_gimg_frac": round(36 * npx / 1e6 / t_bi / peak, 3),
               "fwd_bwd_frac": round(48 * npx / 1e6 / (t_f + t_b) / peak, 3)}
        if S == 1:   # the single-step backward of the same op, for comparison
            o1 = ops[:, 0].contiguous()
            P1 = P[:, 0].contiguous()
            g1 = torch.zeros((B, 24), device=dev)
            t1 = timed(lambda: _lib.check(L.aisp_pointwise_bwd(img.data_ptr(), g.data_ptr(), P1.data_ptr(), o1.data_ptr(), B, H, W,
                                                               1, g1.data_ptr(), None, sc.data_ptr(), sc.numel(), st), "bwd1"))
            row["single_step_bwd_ms"] = round(t1, 4)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
