"""NLM forward timing (with the d/dh stash, as in training) at the bench workload and at the 4K case.
Run once per layout: AISP_NLM_LAYOUT=1col selects the one-column-per-lane kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from adaptiveisp_b200 import _lib
from adaptiveisp_b200 import functional as AF
from adaptiveisp_b200.synthetic import lod_batch

dev = torch.device("cuda:0")
L = _lib.lib()


def run(B, H, W, letterbox, grad):
    img = lod_batch(B, H, W, seed=1235, device=dev) if letterbox else torch.rand((B, 3, H, W), device=dev)
    out = torch.empty_like(img)
    stash = torch.empty_like(img) if grad else None
    P = torch.zeros((B, AF.PSTRIDE), device=dev)
    P[:, 0] = 0.05
    ops = torch.full((B,), AF.OP_NLM, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def call():
        rc = L.aisp_nlm_fwd(img.data_ptr(), out.data_ptr(), P.data_ptr(), ops.data_ptr(), B, H, W,
                            stash.data_ptr() if grad else None, None, st)
        assert rc == 0, rc
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"layout={os.environ.get('AISP_NLM_LAYOUT', '2col')} B={B} {H}x{W} letterbox={letterbox} grad={grad}: "
          f"{ms:.3f} ms  {B * H * W / ms / 1e6:.2f} Gpx/s  checksum {float(out.double().sum()):.6f}")


run(64, 512, 512, True, True)
run(64, 512, 512, False, True)
run(64, 512, 512, False, False)
run(8, 2160, 3840, False, True)
