"""Touch every kernel once at small, ragged sizes (used under compute-sanitizer)."""
import sys
import torch
sys.path.insert(0, ".")
from adaptiveisp_b200 import functional as AF, replay

dev = torch.device("cuda:0")
torch.manual_seed(0)
for (B, H, W) in [(3, 37, 53), (2, 48, 64), (1, 9, 130)]:
    img = (torch.rand((B, 3, H, W), device=dev) * 1.2 - 0.1)
    g = torch.randn_like(img)
    for op in range(13):
        n = AF.NUM_PARAMS[op]
        p = torch.rand((B, n), device=dev) * 0.6 + 0.5
        if op == AF.OP_CCM:
            p = torch.eye(3, device=dev).reshape(1, 9).repeat(B, 1) + 0.1 * torch.rand((B, 9), device=dev)
        for clip in (True, False):
            x = img.clone().requires_grad_(True)
            pp = p.clone().requires_grad_(True)
            y = AF.apply_filter(x, pp, op, clip)
            (y * g).sum().backward()
    ops = torch.tensor([(i * 5) % 13 for i in range(B)], dtype=torch.int32, device=dev)
    P = torch.rand((B, 24), device=dev) * 0.5 + 0.6
    x = img.clone().requires_grad_(True)
    Pq = P.clone().requires_grad_(True)
    y = AF.apply_ops(x, Pq, ops, True)
    (y * g).sum().backward()
    # filter bank: all 13 ops on the same batch (fused per-pixel bank kernels, compact grids)
    bank_ops = list(range(13))
    Pb = (torch.rand((B, 13, 24), device=dev) * 0.5 + 0.6).requires_grad_(True)
    yb = AF.apply_bank(img, Pb, bank_ops, clip=True)
    (yb * torch.randn_like(yb)).sum().backward()
    # device-side selection (all three modes) + the row-gather backward
    Fsel = 10
    pdf = torch.softmax(torch.randn((B, Fsel), device=dev), dim=1)
    st = torch.zeros((B, 3 + Fsel), device=dev)
    tab = torch.arange(Fsel, dtype=torch.int32, device=dev)
    for mode in (AF.SELECT_SAMPLE, AF.SELECT_ARGMAX, AF.SELECT_FORCED):
        pk = torch.randn((B, Fsel, 24), device=dev, requires_grad=True)
        rows = AF.select_rows(pdf, torch.rand((B, 1), device=dev), st, pk, tab, mode, 3)[0]
        rows.sum().backward()
    steps = [[0, 1, 3, 9, 4][: 1 + b % 5] for b in range(B)]
    params = [[torch.rand(AF.NUM_PARAMS[o]) * 0.5 + 0.5 for o in s] for s in steps]
    plan = replay.plan_pipeline(steps, params, dev)
    out = replay.execute_plan(img, plan, True)
    # round 2: sequence launch set with a high-resolution twin and block means, fused sequence backward
    # (6 stages, ColorFilter), differentiable replay, regressors, value statistics
    hi = torch.rand((B, 3, H + 11, W + 7), device=dev)
    out2, hi2 = replay.execute_plan(img, plan, True, high_res=hi)
    oh, ow = (H // 2 if H % 2 == 0 else H), (W // 4 if W % 4 == 0 else W)
    y3, _, d3 = AF.apply_ops(img.clone().requires_grad_(True), P.clone().requires_grad_(True), ops, True, down_hw=(oh, ow))
    (y3.sum() + d3.sum()).backward()
    seq = torch.tensor([[11, 5, 2, 7, 6, 0]] * B, dtype=torch.int32, device=dev)
    lens = torch.tensor([(6 - b) for b in range(B)], dtype=torch.int32, device=dev)
    Pc = (torch.rand((B, 6, 24), device=dev) * 0.5 + 0.6)
    Pc[:, 2, :9] = torch.eye(3, device=dev).reshape(1, 9) + 0.05
    Pc.requires_grad_(True)
    xc = img.clone().requires_grad_(True)
    (AF.apply_chain(xc, Pc, seq, lens, clip_each=True) * g).sum().backward()
    from adaptiveisp_b200.replay_grad import apply_plan
    Pg = [ph.params.clone().requires_grad_(True) for ph in plan.phases]
    (apply_plan(img.clone().requires_grad_(True), plan, Pg) * g).sum().backward()
    AF.value_stats(AF.block_mean(torch.rand((B, 3, 64, 96), device=dev), (16, 24)))
    # synthetic-RAW noise (in-kernel Philox and injected normals), bare NLM modules incl. NonLocalMeansParam
    from adaptiveisp_b200 import unprocess as U, denoise as D
    U.add_read_and_shot_noise(img.clamp(0, 1), [0.01] * B, [1e-4] * B, gain=[0.2] * B, seed=3)
    U.add_read_and_shot_noise(img.clamp(0, 1), 0.01, 1e-4, z=torch.randn_like(img))
    hh = torch.full((B, 1, 1, 1), 0.3, device=dev, requires_grad=True)
    (D.NonLocalMeansGray(11, 5)(img, hh).sum() + D.NonLocalMeans(11, 5)(img, hh).sum()).backward()
    if H >= 5 and W >= 5:
        pm = D.NonLocalMeansParam(0.3, search_window_size=7).to(dev)
        pm(img).sum().backward()
from adaptiveisp_b200 import filters as Fm
from adaptiveisp_b200.config import make_cfg
cfg = make_cfg(feature_extractor_dims=64, fc1_size=16)
mods = [c(cfg, predict=True).to(dev) for c in cfg.filters]
Fm.BankPredictor(mods)(torch.randn((5, 64), device=dev)).sum().backward()
torch.cuda.synchronize()
print("sanity_all_kernels: ok", float(out.mean()))
