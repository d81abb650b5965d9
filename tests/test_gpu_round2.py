"""GPU parity tests added in round 2: BASELINE's own configurations at their own sizes, per-element
gradient error, NaN / inf propagation, the drop-in surface that had no GPU coverage, the USM frame
rule, and the fused sequence backward (up to 6 stages, ColorFilter included).

Tolerances (north star): outputs max-abs <= 1e-5 on [0,1] images; parameter gradients <= 1e-4
RELATIVE PER ELEMENT, where an element smaller than GRAD_FLOOR (1 %) of its tensor's largest entry
is measured against that floor (below it the reference's own fp32 summation noise dominates).
"""
import numpy as np
import pytest
import torch

from oracle import isp_oracle as O
from tests import cases
from tests.test_gpu_parity import OUT_ATOL, GRAD_RTOL, cls_for, out_err, rel_err, _bank_params  # noqa: F401

pytestmark = pytest.mark.gpu

GRAD_FLOOR = 1e-2


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def F(dev):
    from adaptiveisp_b200 import _lib, filters
    _lib.lib()
    return filters


@pytest.fixture(scope="module")
def cfg():
    from adaptiveisp_b200.config import make_cfg
    return make_cfg()


def grad64(op, img, p, g, clip=True):
    """The reference formula of one filter evaluated in fp64 (oracle on double tensors) -> d/d param."""
    p64 = p.double().clone().requires_grad_(True)
    y = (O.forward if clip else O.run)(op, img.double(), p64)
    (y * g.double()).sum().backward()
    return p64.grad.reshape(-1).numpy()


def elem_err(a, b, floor=GRAD_FLOOR):
    """Worst per-element relative error |a-b| / max(|b|, floor * max|b|) and where it occurs."""
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    den = np.maximum(np.abs(b), floor * max(np.abs(b).max(), 1e-30))
    e = np.abs(a - b) / den
    i = int(np.argmax(e))
    return float(e[i]), i


def assert_grad(a, b, tol, tag):
    e, i = elem_err(a, b)
    assert e <= tol, f"{tag}: worst component {i}: got {np.asarray(a).reshape(-1)[i]:.8g} " \
                     f"want {np.asarray(b).reshape(-1)[i]:.8g} (per-element rel err {e:.3g} > {tol})"


class Checks:
    """Collects every violated bound of a test and reports them together (one GPU run shows them all)."""

    def __init__(self):
        self.bad = []

    def out(self, got, ref, tol, tag):
        e = out_err(got, ref)
        if not e <= tol:
            self.bad.append(f"{tag}: output err {e:.3g} > {tol}")

    def grad(self, a, b, tol, tag, ref64=None):
        """Per-element bound against the reference's fp32 autograd; where that fails and an fp64 evaluation
        of the same formula is supplied (`ref64`, a callable, evaluated lazily), the CUDA value must be at
        least as close to it as the reference's own fp32 value is (its summation noise is then the cause)."""
        e, i = elem_err(a, b)
        if e <= tol:
            return
        msg = (f"{tag}: component {i} got {np.asarray(a).reshape(-1)[i]:.8g} want "
               f"{np.asarray(b).reshape(-1)[i]:.8g} (per-element rel err {e:.3g} > {tol})")
        if ref64 is not None:
            r64 = np.asarray(ref64(), np.float64).reshape(-1)
            a64, b64 = np.asarray(a, np.float64).reshape(-1), np.asarray(b, np.float64).reshape(-1)
            den = np.maximum(np.abs(r64), GRAD_FLOOR * np.abs(r64).max())
            e_cuda, e_ref = np.abs(a64 - r64) / den, np.abs(b64 - r64) / den
            if np.all(e_cuda <= np.maximum(tol, 1.5 * e_ref)):
                return
            j = int(np.argmax(e_cuda - np.maximum(tol, 1.5 * e_ref)))
            msg += f"; vs fp64: cuda err {e_cuda[j]:.3g}, reference-fp32 err {e_ref[j]:.3g} at component {j}"
        self.bad.append(msg)

    def norm(self, a, b, tol, tag):
        e = rel_err(a, b)
        if not e <= tol:
            self.bad.append(f"{tag}: norm-wise rel err {e:.3g} > {tol}")

    def done(self):
        assert not self.bad, "\n".join(self.bad)


# ----------------------------------------------------------------------------------------------
# 1. BASELINE configs[1] at its own size: 64 x 3 x 512 x 512, ten filters, forward + backward
# ----------------------------------------------------------------------------------------------
def test_config2_full_size_bank_vs_oracle(dev):
    """The bench workload itself: FilterBank stack and all ten parameter-gradient rows at B=64, 512x512
    (262k-pixel reductions: 64 scratch rows per sample -> fp64 finalize), compared with the CPU oracle
    on two images of the batch (NLM on one: ~8k ATen ops at 512x512 per call)."""
    from adaptiveisp_b200 import functional as AF
    ops = cases.AGENT_OPS
    B, H, W = 64, 512, 512
    Fn = len(ops)
    img = cases.lod_batch(B, H, W, seed=1235)
    P, plist = _bank_params(ops, B, seed=500)
    gen = torch.Generator(device=dev).manual_seed(77)
    g = torch.randn((B, Fn, 3, H, W), device=dev, generator=gen)
    g[:, ops.index(O.OP_NLM)].abs_()       # d/dh under a random-sign g: the reference's own fp32 sum is noise
    Pd = P.to(dev).requires_grad_(True)
    y = AF.apply_bank(img.to(dev), Pd, ops, clip=True)
    (y * g).sum().backward()
    torch.cuda.synchronize()
    checked, chk = 0, Checks()
    for b in (5, 40):
        for f, op in enumerate(ops):
            if op == O.OP_NLM and b != 5:
                continue
            pc = plist[f][b:b + 1].clone().requires_grad_(True)
            yc = O.forward(op, img[b:b + 1], pc)
            (yc * g[b:b + 1, f].cpu()).sum().backward()
            tag = f"image {b} {O.OP_NAMES[op]}"
            chk.out(y[b:b + 1, f].detach().cpu().numpy(), yc.detach().numpy(), OUT_ATOL, tag)
            n = O.OP_NPARAMS[op]
            chk.grad(Pd.grad[b, f, :n].cpu().numpy(), pc.grad.reshape(-1).numpy(), GRAD_RTOL, tag,
                     ref64=(None if op == O.OP_NLM else
                            (lambda op=op, b=b, f=f: grad64(op, img[b:b + 1], plist[f][b:b + 1], g[b:b + 1, f].cpu()))))
            checked += 1
    chk.done()
    assert checked == 19


def test_config2_full_size_agent_select_vs_oracle(dev):
    """Agent semantics at the bench size: ONE selected filter per sample (heterogeneous launch, B=64,
    512x512), outputs and parameter gradients of a sample per filter against the oracle."""
    from adaptiveisp_b200 import functional as AF
    B, H, W = 64, 512, 512
    rng = np.random.RandomState(3)
    ops = [cases.AGENT_OPS[i % 10] for i in range(B)]
    rng.shuffle(ops)
    img = cases.lod_batch(B, H, W, seed=1240)
    rows, plist = [], []
    for b, op in enumerate(ops):
        _, p = cases.params_for(op, 1, seed=900 + b)
        plist.append(p)
        rows.append(AF.pack_params(p, O.OP_NPARAMS[op]))
    Pd = torch.cat(rows, 0).to(dev).requires_grad_(True)
    gen = torch.Generator(device=dev).manual_seed(78)
    g = torch.randn((B, 3, H, W), device=dev, generator=gen)
    for b, op in enumerate(ops):
        if op == O.OP_NLM:
            g[b].abs_()
    y = AF.apply_ops(img.to(dev), Pd, torch.tensor(ops, dtype=torch.int32, device=dev), clip=True)
    (y * g).sum().backward()
    torch.cuda.synchronize()
    seen, chk = set(), Checks()
    for b, op in enumerate(ops):
        if op in seen:
            continue
        seen.add(op)
        pc = plist[b].clone().requires_grad_(True)
        yc = O.forward(op, img[b:b + 1], pc)
        (yc * g[b:b + 1].cpu()).sum().backward()
        tag = f"sample {b} {O.OP_NAMES[op]}"
        chk.out(y[b:b + 1].detach().cpu().numpy(), yc.detach().numpy(), OUT_ATOL, tag)
        n = O.OP_NPARAMS[op]
        chk.grad(Pd.grad[b, :n].cpu().numpy(), pc.grad.reshape(-1).numpy(), GRAD_RTOL, tag,
                 ref64=(None if op == O.OP_NLM else
                        (lambda op=op, b=b: grad64(op, img[b:b + 1], plist[b], g[b:b + 1].cpu()))))
    chk.done()
    assert len(seen) == 10


# ----------------------------------------------------------------------------------------------
# 2. BASELINE configs[3] (4K, desaturation / NLM / USM) backward
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("op", [O.OP_WNB, O.OP_NLM, O.OP_USM])
def test_config4_4k_backward_on_windows(dev, op):
    """d/dp (BW), d/dh (NLM), d/dsigma and d/damount (USM) at 3840x2160.  A parameter gradient is a sum
    over the whole frame, which the CPU oracle cannot differentiate at 4K (7.6k image-sized autograd
    temporaries for NLM); so the upstream gradient is supported on a window, the oracle runs on the
    crop around it (window + the stencil's dependency radius) and must give the same number."""
    from adaptiveisp_b200 import functional as AF
    B, H, W = 2, 2160, 3840
    img = cases.lod_batch(B, H, W, seed=1237, letterbox=False)
    m, cs = 12, 96                                   # margin >= 7 (NLM) / 2 (USM); crop size
    wins = [(700, 1500), (2160 - cs, 3840 - cs)]     # an interior window and the bottom-right corner
    pv = {O.OP_WNB: [[0.35], [0.55]], O.OP_NLM: [[0.25], [0.55]], O.OP_USM: [[0.9, 1.2], [1.4, 0.7]]}[op]
    p = torch.tensor(pv)
    g = torch.zeros((B, 3, H, W))
    gw = cases.grad_out((B, 3, cs - 2 * m, cs - 2 * m), 41)
    if op == O.OP_NLM:
        gw = gw.abs()
    for b, (y0, x0) in enumerate(wins):
        g[b, :, y0 + m:y0 + cs - m, x0 + m:x0 + cs - m] = gw[b]
    Pd = AF.pack_params(p, O.OP_NPARAMS[op]).to(dev).requires_grad_(True)
    xd = img.to(dev).requires_grad_(op != O.OP_NLM)          # the NLM image gradient is the slow rare path
    y = AF.apply_ops(xd, Pd, op, clip=True)
    (y * g.to(dev)).sum().backward()
    torch.cuda.synchronize()
    for b, (y0, x0) in enumerate(wins):
        # the corner window sits on the frame: take the crop so that the frame rule is the image's own
        # (USM reflects, the 3x3 keeps the border); NLM wraps circularly, which no crop reproduces, so its
        # corner window is moved one radius inside
        if op == O.OP_NLM and b == 1:
            continue
        xc = img[b:b + 1, :, y0:y0 + cs, x0:x0 + cs].clone().requires_grad_(True)
        pc = p[b:b + 1].clone().requires_grad_(True)
        yc = O.forward(op, xc, pc)
        gc = g[b:b + 1, :, y0:y0 + cs, x0:x0 + cs]
        (yc * gc).sum().backward()
        tag = f"{O.OP_NAMES[op]} window {b}"
        got = y[b:b + 1, :, y0 + m:y0 + cs - m, x0 + m:x0 + cs - m].detach().cpu().numpy()
        assert out_err(got, yc[:, :, m:cs - m, m:cs - m].detach().numpy()) <= OUT_ATOL, tag
        n = O.OP_NPARAMS[op]
        assert_grad(Pd.grad[b, :n].cpu().numpy(), pc.grad.reshape(-1).numpy(), 2 * GRAD_RTOL, tag)
        if op != O.OP_NLM:
            gi = xd.grad[b:b + 1, :, y0 + m:y0 + cs - m, x0 + m:x0 + cs - m].cpu().numpy()
            assert rel_err(gi, xc.grad[:, :, m:cs - m, m:cs - m].numpy()) <= GRAD_RTOL, tag


# ----------------------------------------------------------------------------------------------
# 3. drop-in surface without GPU coverage so far: ToneFilterV2.process, run_v2, predict_param
# ----------------------------------------------------------------------------------------------
def test_tone_v2_flat_params_run_v2_predict_param(F, cfg, dev):
    """isp/filters.py:365-387 (flat [B,8] curve), :141-152 (run_v2: parameter without the batch dim),
    :154-159 (predict_param: features -> regressor -> unclipped process)."""
    B, H, W = 3, 40, 52
    img = cases.edge_image(B, H, W, seed=5)
    _, p5 = cases.params_for(O.OP_TONE, B, seed=5)            # [B,8,1,1,1]
    flat = p5.reshape(B, 8)
    ref = O.run(O.OP_TONE, img, p5)
    v2 = F.ToneFilterV2(cfg, predict=True).to(dev)
    pd = flat.to(dev).requires_grad_(True)
    got = v2.process(img.to(dev), pd)                         # flat layout, as the reference's V2 expects
    assert out_err(got.detach().cpu().numpy(), ref.numpy()) <= OUT_ATOL
    g = cases.grad_out(img.shape, 5)
    (got * g.to(dev)).sum().backward()
    pc = p5.clone().requires_grad_(True)
    (O.run(O.OP_TONE, img, pc) * g).sum().backward()
    assert_grad(pd.grad.cpu().numpy(), pc.grad.reshape(B, 8).numpy(), GRAD_RTOL, "ToneFilterV2.process d/dp")
    # run_v2: one parameter row for a single image
    one = v2.run_v2(img[:1].to(dev), flat[0].to(dev))
    assert out_err(one.cpu().numpy(), ref[:1].numpy()) <= OUT_ATOL
    e = F.ExposureFilter(cfg).to(dev)
    assert out_err(e.run_v2(img[:1].to(dev), torch.tensor([0.7], device=dev)).cpu().numpy(),
                   O.run(O.OP_EXPOSURE, img[:1], torch.tensor([[0.7]])).numpy()) <= OUT_ATOL
    # predict_param: features through the module's own FC layers, no clip
    torch.manual_seed(4)
    for cls in (F.GammaFilter, F.CCMFilter, F.SharpenFilter, F.ToneFilter):
        flt = cls(cfg, predict=True).to(dev)
        feats = torch.randn((B, cfg.feature_extractor_dims), device=dev) * 0.05
        out = flt.predict_param(img.to(dev), feats)
        with torch.no_grad():
            raw, _ = flt.extract_parameters(feats)
            p = flt.filter_param_regressor(raw).cpu()
        assert out_err(out.detach().cpu().numpy(), O.run(flt.OP, img, p).numpy()) <= OUT_ATOL, cls.__name__
        assert flt.mask.shape == (1, 1, 1, 1)


# ----------------------------------------------------------------------------------------------
# 4. non-finite pixels propagate as in ATen (so that the guard of train.py:374 fires on the same batches)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("op", cases.POINTWISE_OPS + [O.OP_SHARPEN, O.OP_USM])
@pytest.mark.parametrize("clip", [True, False])
def test_nonfinite_inputs_propagate_like_the_reference(F, cfg, dev, op, clip):
    """NaN and +-inf planted in single channels: the set of non-finite output values must be the
    reference's, element for element (torch.clamp / maximum / max(dim) keep NaN, and the wrapper's
    lerp `0 * img + process` turns an inf pixel into NaN), and every finite value must still match.
    The stencil filters go through a library convolution in the reference (oneDNN on the CPU, cuDNN on a
    GPU), whose treatment of inf inside a window is the library's own; there only NaN is planted for the
    exact comparison and inf is checked as "the batch is flagged non-finite on both sides"."""
    B, H, W = 2, 12, 16
    img = cases.edge_image(B, H, W, seed=13, in_range=True)
    nan, inf = float("nan"), float("inf")
    stencil = op in cases.STENCIL_OPS
    img[0, 0, 3, 3] = nan
    img[0, 1, 3, 6] = nan
    img[0, 2, 5, 9] = nan
    img[1, :, 9, 9] = nan
    if not stencil:
        img[1, 0, 4, 4] = inf
        img[1, 1, 6, 7] = -inf
        img[1, 2, 8, 2] = inf
    _, p = cases.params_for(op, B, seed=13)
    ref = (O.forward if clip else O.run)(op, img, p)
    flt = cls_for(F, op)(cfg).to(dev)
    run = (lambda x: flt(x.to(dev), specified_parameter=p.to(dev))[0].cpu()) if clip else (lambda x: flt.run(x.to(dev), p.to(dev)).cpu())
    got = run(img)
    assert torch.equal(torch.isnan(got), torch.isnan(ref)), O.OP_NAMES[op]
    assert torch.equal(torch.isposinf(got), torch.isposinf(ref)) and torch.equal(torch.isneginf(got), torch.isneginf(ref))
    fin = torch.isfinite(ref)
    assert out_err(got[fin].numpy(), ref[fin].numpy()) <= OUT_ATOL, O.OP_NAMES[op]
    assert int(torch.isnan(ref).sum()) >= 4      # the case really exercises the propagation
    if stencil:
        img2 = cases.edge_image(B, H, W, seed=14, in_range=True)
        img2[1, 2, 6, 5] = inf
        ref2, got2 = (O.forward if clip else O.run)(op, img2, p), run(img2)
        assert not bool(torch.isfinite(ref2).all()) and not bool(torch.isfinite(got2).all())
        assert bool(torch.isfinite(got2[0]).all()) and bool(torch.isfinite(ref2[0]).all())


def test_nan_batch_trips_the_pool_refill_guard(F, cfg, dev):
    """train.py:374: `torch.isnan(retouch).any() or torch.isinf(retouch).any()` must be True for exactly
    the batches for which the reference says so -- a NaN made by a singular CCM at step t and fed to
    Tone / Gamma / S+ at step t+1 must survive."""
    from adaptiveisp_b200 import functional as AF
    img = cases.edge_image(2, 16, 16, seed=2, in_range=True)
    p_ccm = torch.tensor([[1.0, -0.5, -0.5, 0.2, 0.5, 0.3, 0.1, 0.1, 0.8]]).repeat(2, 1)   # row 0 sums to 0
    step1_ref = O.forward(O.OP_CCM, img, p_ccm)
    step1 = cls_for(F, O.OP_CCM)(cfg).to(dev)(img.to(dev), specified_parameter=p_ccm.to(dev))[0]
    assert torch.equal(torch.isnan(step1.cpu()), torch.isnan(step1_ref))
    assert bool(torch.isnan(step1_ref).any())
    for op in (O.OP_TONE, O.OP_GAMMA, O.OP_SATPLUS, O.OP_CONTRAST, O.OP_NLM):
        _, p = cases.params_for(op, 2, seed=3)
        ref = O.forward(op, step1_ref, p)
        got = cls_for(F, op)(cfg).to(dev)(step1, specified_parameter=p.to(dev))[0].cpu()
        bad_ref = bool(torch.isnan(ref).any() or torch.isinf(ref).any())
        bad_got = bool(torch.isnan(got).any() or torch.isinf(got).any())
        assert bad_ref and bad_got, O.OP_NAMES[op]
        assert torch.equal(torch.isnan(got), torch.isnan(ref)), O.OP_NAMES[op]
        down = AF.block_mean(got.to(dev), (4, 4))
        assert not bool(AF.image_stats(down)[1].all())           # the one-pass guard sees it too


# ----------------------------------------------------------------------------------------------
# 5. USM: a tile whose 2-px halo leaves the image although the tile itself does not touch the frame
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", [(1, 33, 260), (1, 49, 512), (2, 17, 384), (1, 34, 260), (1, 35, 516),
                                  # one tile holding both frames of an axis, images smaller than the halo box, W % 4 != 0
                                  (1, 3, 4), (1, 5, 8), (2, 18, 132), (1, 16, 128), (1, 4, 128), (1, 7, 9), (1, 19, 131)])
def test_usm_halo_leaves_image_on_tma_path(F, cfg, dev, size):
    """H % 16 == 1 with W % 4 == 0 and W > 256: the last-but-one tile row ends one row short of the frame
    while its halo row y0 + 17 == H is outside; reflect padding (isp/sharpen.py:76-78) must apply there
    (the TMA path zero-fills out-of-bounds elements).  Forward, d/dsigma, d/damount and d/d img."""
    B, H, W = size
    flt = F.SharpenUSMFilter(cfg).to(dev)
    img = cases.edge_image(B, H, W, seed=H, in_range=True)
    p = torch.tensor([[1.1, 1.5], [0.6, 0.8]])[:B]
    g = cases.grad_out(img.shape, H)
    xc, pc = img.clone().requires_grad_(True), p.clone().requires_grad_(True)
    yc = O.forward(O.OP_USM, xc, pc)
    (yc * g).sum().backward()
    xd, pd = img.to(dev).requires_grad_(True), p.to(dev).requires_grad_(True)
    yd = flt(xd, specified_parameter=pd)[0]
    (yd * g.to(dev)).sum().backward()
    err = (yd.detach().cpu() - yc.detach()).abs()
    assert float(err.max()) <= OUT_ATOL, f"worst row {int(err.amax(dim=(0, 1, 3)).argmax())} of {H}"
    assert_grad(pd.grad.cpu().numpy(), pc.grad.numpy(), GRAD_RTOL, "USM d/d(sigma, amount)")
    assert rel_err(xd.grad.cpu().numpy(), xc.grad.numpy()) <= GRAD_RTOL


@pytest.mark.parametrize("size", [(1, 3, 4), (1, 5, 8), (2, 18, 132), (1, 16, 128), (1, 33, 260), (1, 7, 9), (2, 34, 516)])
@pytest.mark.parametrize("op", ["SHARPEN", "SHARPEN_V2"])
def test_sharpen_3x3_frame_and_small_sizes(dev, op, size):
    """3x3 filters on the packed-pair kernels: frame pixels pass x through (isp/sharpen.py:105-142), tiles that hold
    both frames of an axis, images narrower than a tile, the scalar (W % 4 != 0) path; forward, d/df, d/d img
    (transposed stencil with the frame pixels' gradient masked out of the blur)."""
    from adaptiveisp_b200 import functional as AF
    opc = getattr(O, "OP_" + op)
    B, H, W = size
    img = cases.edge_image(B, H, W, seed=H + W, in_range=True)
    p = torch.tensor([[1.7], [0.4]])[:B]
    g = cases.grad_out(img.shape, W)
    # knife-edge pixels: on the k/255 lattice of sample 0 the pre-clip value can be EXACTLY 0 or 1 in exact arithmetic
    # (e.g. 2.7 x = 1.7 blur); which side of the clamp it lands on in fp32 depends on the summation order of the 3x3
    # convolution (oneDNN's, in the reference), so the gradient mask of such a pixel is not defined by the arithmetic.
    # They are found in float64 and taken out of the upstream gradient.
    x64, f64 = img.double(), p.double()[:, :, None, None]
    b64 = O._blur3x3_keep_border(x64)
    y64 = x64 * f64 + b64 * (1.0 - f64) if op == "SHARPEN" else x64 + (x64 - b64) * f64
    g = g * ((y64.abs() > 1e-6) & ((y64 - 1.0).abs() > 1e-6)).float()
    xc, pc = img.clone().requires_grad_(True), p.clone().requires_grad_(True)
    yc = O.forward(opc, xc, pc)
    (yc * g).sum().backward()
    xd, pd = img.to(dev).requires_grad_(True), p.to(dev).requires_grad_(True)
    yd = AF.apply_filter(xd, pd, opc, True)
    (yd * g.to(dev)).sum().backward()
    assert float((yd.detach().cpu() - yc.detach()).abs().max()) <= OUT_ATOL
    assert_grad(pd.grad.cpu().numpy(), pc.grad.numpy(), GRAD_RTOL, op + " d/df")
    assert rel_err(xd.grad.cpu().numpy(), xc.grad.numpy()) <= GRAD_RTOL


# ----------------------------------------------------------------------------------------------
# 6. fused sequence backward, second generation: up to 6 stages, ColorFilter, strictness, identity
# ----------------------------------------------------------------------------------------------
def _chain_case(B, S, H, W, seed, pool, lens=None, fixed_first=None):
    rng = np.random.RandomState(seed)
    ops = rng.choice(pool, size=(B, S))
    if fixed_first is not None:
        ops[0, :len(fixed_first)] = fixed_first
    if lens is None:
        lens = rng.randint(1, S + 1, size=B)
        lens[0] = S if fixed_first is None else len(fixed_first)
    img = cases.edge_image(B, H, W, seed=seed, in_range=True)
    P = torch.zeros((B, S, 24))
    plist = [[None] * S for _ in range(B)]
    for b in range(B):
        for k in range(S):
            _, p = cases.params_for(int(ops[b, k]), 1, seed=70 + 5 * b + k + seed)
            plist[b][k] = p
            P[b, k, :O.OP_NPARAMS[int(ops[b, k])]] = p.reshape(-1)
    return ops, np.asarray(lens), img, P, plist


@pytest.mark.parametrize("clip_each", [True, False])
@pytest.mark.parametrize("S,size", [(5, (6, 40, 52)), (6, (4, 33, 47)), (6, (3, 64, 64)), (2, (3, 17, 23))])
def test_fused_chain_up_to_six_stages_with_color(dev, clip_each, S, size):
    """Sequences of up to AISP_MAX_CHAIN_BWD = 6 per-pixel filters (BASELINE configs[4] draws up to 5),
    ColorFilter included, forward in one pass and backward in one pass: every stage's parameter
    gradients and the image gradient against autograd through the CPU oracle.  (33x47 / 17x23: the
    scalar cp.async path with a ragged tail; 64x64: exactly one 4096-pixel chunk.)"""
    from adaptiveisp_b200 import functional as AF
    B, H, W = size
    ops, lens, img, P, plist = _chain_case(B, S, H, W, 31 + S, cases.POINTWISE_OPS,
                                           fixed_first=[O.OP_COLOR, O.OP_TONE, O.OP_CCM, O.OP_COLOR][:min(S, 4)])
    g = cases.grad_out(img.shape, 31)
    xd = img.to(dev).requires_grad_(True)
    Pd = P.to(dev).requires_grad_(True)
    y = AF.apply_chain(xd, Pd, torch.tensor(ops, dtype=torch.int32, device=dev),
                       torch.tensor(lens, dtype=torch.int32, device=dev), clip_each=clip_each)
    (y * g.to(dev)).sum().backward()
    # rounding compounds over the stages on BOTH sides (a gamma with p < 1 amplifies an upstream 1e-7 by
    # up to p * 0.001^(p-1) ~ 30): the per-filter bars (1e-5 / 1e-4) are widened by the chain length
    out_tol, grad_tol = 1e-5 * max(S, 3), 1e-4 * max(S, 3)
    chk = Checks()
    for b in range(B):
        n = int(lens[b])
        xc = img[b:b + 1].clone().requires_grad_(True)
        pcs = [plist[b][k].clone().requires_grad_(True) for k in range(n)]
        yc = O.chain([int(o) for o in ops[b, :n]], xc, pcs, clip_each)
        (yc * g[b:b + 1]).sum().backward()
        tag = f"sample {b} {[O.OP_NAMES[int(o)] for o in ops[b, :n]]}"
        chk.out(y[b:b + 1].detach().cpu().numpy(), yc.detach().numpy(), out_tol, tag)
        chk.norm(xd.grad[b:b + 1].cpu().numpy(), xc.grad.numpy(), grad_tol, tag + " d/d img")
        for k in range(n):
            m = O.OP_NPARAMS[int(ops[b, k])]
            ref = pcs[k].grad.reshape(-1).numpy()
            if np.abs(ref).max() > 1e-6:
                chk.grad(Pd.grad[b, k, :m].cpu().numpy(), ref, grad_tol, f"{tag} stage {k}")
        if n < S:
            assert float(Pd.grad[b, n:].abs().max()) == 0.0
    chk.done()


def test_fused_chain_full_size_e_g_wb_ccm(dev):
    """The chain of isp/filters.py:753-815 (E -> G -> WB -> CCM, the reference's own parameters) at the
    bench size 64 x 512 x 512: fused forward and fused backward against the oracle on two images."""
    from adaptiveisp_b200 import functional as AF
    B, H, W = 64, 512, 512
    img = cases.lod_batch(B, H, W, seed=1234)
    ops = [O.OP_EXPOSURE, O.OP_GAMMA, O.OP_WB, O.OP_CCM]
    P = torch.zeros((B, 4, 24))
    P[:, 0, 0] = 0.09012079
    P[:, 1, 0] = 0.38566995
    P[:, 2, :3] = torch.tensor([2.4052505, 1.2233436, 1.8800205])
    P[:, 3, :9] = torch.tensor([1.6, -0.4, -0.2, -0.3, 1.5, -0.2, -0.1, -0.5, 1.6])
    P[:, 0, 0] += torch.linspace(-0.5, 1.5, B)               # per-sample exposure: dark and clipped frames
    gen = torch.Generator(device=dev).manual_seed(5)
    g = torch.randn((B, 3, H, W), device=dev, generator=gen)
    Pd = P.to(dev).requires_grad_(True)
    y = AF.apply_chain(img.to(dev), Pd, torch.tensor([ops] * B, dtype=torch.int32, device=dev), clip_each=True)
    (y * g).sum().backward()
    torch.cuda.synchronize()
    chk = Checks()
    for b in (3, 60):
        pcs = [P[b:b + 1, k, :O.OP_NPARAMS[op]].clone().requires_grad_(True) for k, op in enumerate(ops)]
        yc = O.chain(ops, img[b:b + 1], pcs, True)
        (yc * g[b:b + 1].cpu()).sum().backward()
        chk.out(y[b:b + 1].detach().cpu().numpy(), yc.detach().numpy(), 4e-5, f"image {b}")   # four stages
        for k, op in enumerate(ops):
            chk.grad(Pd.grad[b, k, :O.OP_NPARAMS[op]].cpu().numpy(), pcs[k].grad.reshape(-1).numpy(), 4e-4,
                     f"image {b} stage {k} {O.OP_NAMES[op]}")
    chk.done()


def test_fused_chain_identity_none_and_strictness(dev):
    """seq_len == 0 is the identity (gradient passes through), AISP_OP_NONE gives a zero image and zero
    gradients, and -- the ops live on the device, nothing is read back -- a stencil op or an unknown
    code inside a fused sequence poisons that sample with NaN instead of leaving stale memory."""
    from adaptiveisp_b200 import AispError, functional as AF
    B, S, H, W = 5, 3, 20, 24
    img = cases.edge_image(B, H, W, seed=3, in_range=True)
    ops = torch.tensor([[O.OP_GAMMA, O.OP_TONE, O.OP_WB],
                        [O.OP_GAMMA, O.OP_TONE, O.OP_WB],
                        [-1, O.OP_GAMMA, O.OP_GAMMA],
                        [O.OP_SHARPEN, O.OP_GAMMA, O.OP_GAMMA],
                        [O.OP_EXPOSURE, O.OP_NLM, O.OP_GAMMA]], dtype=torch.int32, device=dev)
    lens = torch.tensor([3, 0, 3, 3, 3], dtype=torch.int32, device=dev)
    P = torch.full((B, S, 24), 0.8, device=dev).requires_grad_(True)
    x = img.to(dev).requires_grad_(True)
    g = cases.grad_out(img.shape, 3).to(dev)
    y = AF.apply_chain(x, P, ops, lens, clip_each=True)
    (torch.nan_to_num(y) * g).sum().backward()
    assert torch.equal(y[1], x[1].detach()) and torch.equal(x.grad[1], g[1])            # identity
    assert float(P.grad[1].abs().max()) == 0.0
    assert float(y[2].abs().max()) == 0.0 and float(x.grad[2].abs().max()) == 0.0       # AISP_OP_NONE
    assert float(P.grad[2].abs().max()) == 0.0
    for b in (3, 4):                                                                     # stencil op: NaN, not garbage
        assert bool(torch.isnan(y[b]).all()) and bool(torch.isnan(x.grad[b]).all())
        assert bool(torch.isnan(P.grad[b]).all())
    assert bool(torch.isfinite(y[0]).all()) and bool(torch.isfinite(P.grad[0, :, 0]).all())
    with pytest.raises(AispError):                                                       # 7 > AISP_MAX_CHAIN_BWD with grads
        AF.apply_chain(x, torch.zeros((B, 7, 24), device=dev, requires_grad=True),
                       torch.zeros((B, 7), dtype=torch.int32, device=dev))
    y7 = AF.apply_chain(x.detach(), torch.zeros((B, 7, 24), device=dev), torch.zeros((B, 7), dtype=torch.int32, device=dev))
    assert y7.shape == x.shape                                                           # forward-only: up to 8 steps


# ----------------------------------------------------------------------------------------------
# 7. sequence launch set (aisp_sequence_fwd): one stencil step anywhere in a sequence, one pass over HBM
# ----------------------------------------------------------------------------------------------
def _seq_case(seqs, H, W, seed):
    B = len(seqs)
    S = max(len(s) for s in seqs)
    img = cases.edge_image(B, H, W, seed=seed, in_range=True)
    ops = torch.zeros((B, S), dtype=torch.int32)
    lens = torch.tensor([len(s) for s in seqs], dtype=torch.int32)
    P = torch.zeros((B, S, 24))
    plist = []
    for b, seq in enumerate(seqs):
        row = []
        for k, op in enumerate(seq):
            _, p = cases.params_for(op, 1, seed=seed + 13 * b + k)
            if op == O.OP_NLM:
                p = torch.tensor([[0.15 + 0.1 * b]])
            row.append(p)
            ops[b, k] = op
            P[b, k, :O.OP_NPARAMS[op]] = p.reshape(-1)
        plist.append(row)
    return img, ops, lens, P, plist


SEQS = [
    [O.OP_EXPOSURE, O.OP_GAMMA, O.OP_WB, O.OP_CCM, O.OP_SHARPEN],      # isp/filters.py:753-815 in ONE launch
    [O.OP_EXPOSURE, O.OP_GAMMA, O.OP_WB, O.OP_CCM, O.OP_USM],          # ... and its USM variant
    [O.OP_WNB, O.OP_NLM, O.OP_TONE],                                   # prologue + NLM + epilogue
    [O.OP_SHARPEN_V2, O.OP_CONTRAST, O.OP_SATPLUS],                    # stencil first
    [O.OP_TONE, O.OP_CCM],                                             # no stencil at all
    [O.OP_NLM],                                                        # a lone stencil
    [O.OP_GAMMA, O.OP_COLOR, O.OP_USM, O.OP_EXPOSURE, O.OP_WNB, O.OP_TONE, O.OP_CCM, O.OP_WB],   # 8 steps
    [],                                                                # idle sample: carried through
]


@pytest.mark.parametrize("clip_each", [True, False])
@pytest.mark.parametrize("size", [(40, 56), (33, 47), (48, 264)])
def test_sequence_launch_set_matches_oracle(dev, clip_each, size):
    """Heterogeneous per-sample sequences, each with at most one stencil step somewhere in it, run by ONE
    launch set; (33,47): scalar / cp.async paths; (48,264): W % 4 == 0 and W > 256 -> TMA tiles, interior
    and frame, a USM tile whose halo leaves the image."""
    from adaptiveisp_b200 import functional as AF
    H, W = size
    img, ops, lens, P, plist = _seq_case(SEQS, H, W, seed=5 + H)
    got, _, _ = AF.sequence_forward(img.to(dev), P.to(dev), ops.to(dev), lens.to(dev), clip_each=clip_each)
    got = got.cpu()
    chk = Checks()
    for b, seq in enumerate(SEQS):
        ref = O.chain(seq, img[b:b + 1], plist[b], clip_each) if seq else img[b:b + 1]
        chk.out(got[b:b + 1].numpy(), ref.numpy(), 1e-5 * max(len(seq), 3), f"sample {b} {[O.OP_NAMES[o] for o in seq]}")
    chk.done()


def test_sequence_second_stencil_ends_the_sequence_and_planner_splits(dev):
    """Two stencil steps cannot share one pass (the second needs the first's output at neighbouring
    pixels): the kernel stops a sequence in front of the second one, and replay.plan_pipeline splits such
    pipelines into phases -- BASELINE configs[3]'s BW -> NLM -> USM becomes [BW -> NLM], [USM]."""
    from adaptiveisp_b200 import functional as AF, replay
    seqs = [[O.OP_WNB, O.OP_NLM, O.OP_USM, O.OP_GAMMA], [O.OP_SHARPEN, O.OP_SHARPEN]]
    img, ops, lens, P, plist = _seq_case(seqs, 36, 44, seed=9)
    got, _, _ = AF.sequence_forward(img.to(dev), P.to(dev), ops.to(dev), lens.to(dev))
    assert out_err(got[0:1].cpu().numpy(), O.chain(seqs[0][:2], img[0:1], plist[0][:2], True).numpy()) <= 3e-5
    assert out_err(got[1:2].cpu().numpy(), O.chain(seqs[1][:1], img[1:2], plist[1][:1], True).numpy()) <= OUT_ATOL
    plan = replay.plan_pipeline(seqs, plist, dev)
    assert len(plan.phases) == 2
    full = replay.execute_plan(img.to(dev), plan).cpu()
    for b in range(2):
        assert out_err(full[b:b + 1].numpy(), O.chain(seqs[b], img[b:b + 1], plist[b], True).numpy()) <= 4e-5, b


def test_high_res_twin_in_the_same_launch_set(dev):
    """isp/filters.py:116-122 / agent.py:155-157 / train.py:541: the sequences and parameters of the
    low-resolution batch applied to a full-size twin (odd, unaligned shape: its own scalar launch; aligned
    shape: extra tiles of the same launch)."""
    from adaptiveisp_b200 import functional as AF
    seqs = [[O.OP_GAMMA], [O.OP_SHARPEN], [O.OP_NLM], [O.OP_EXPOSURE, O.OP_USM, O.OP_TONE], [O.OP_SATPLUS, O.OP_CCM]]
    img, ops, lens, P, plist = _seq_case(seqs, 64, 64, seed=21)
    for (H2, W2) in [(75, 131), (96, 160)]:
        hi = cases.lod_batch(len(seqs), H2, W2, seed=22, letterbox=False)
        lo_out, hi_out, _ = AF.sequence_forward(img.to(dev), P.to(dev), ops.to(dev), lens.to(dev), high_res=hi.to(dev))
        chk = Checks()
        for b, seq in enumerate(seqs):
            chk.out(lo_out[b:b + 1].cpu().numpy(), O.chain(seq, img[b:b + 1], plist[b], True).numpy(), 3e-5, f"low {b}")
            chk.out(hi_out[b:b + 1].cpu().numpy(), O.chain(seq, hi[b:b + 1], plist[b], True).numpy(), 3e-5, f"high {b} {H2}x{W2}")
        chk.done()


@pytest.mark.parametrize("size,out", [((512, 512), (64, 64)), ((256, 256), (64, 64)), ((96, 160), (32, 32)),
                                      ((64, 64), (64, 64)), ((128, 512), (16, 64))])
def test_block_means_leave_through_the_store_path(dev, size, out):
    """the pooled image that agent.py:97 / value.py:63 compute next comes out of the apply
    itself -- from the per-pixel and sharpen kernels' store path where pooling blocks tile a CTA's work
    (512 -> 64, 256 -> 64), from one masked block-mean pass otherwise (NLM samples, 3x5 blocks)."""
    from adaptiveisp_b200 import functional as AF
    H, W = size
    ops = [O.OP_GAMMA, O.OP_SHARPEN, O.OP_NLM, -1, O.OP_USM, O.OP_CCM, O.OP_SATPLUS]
    B = len(ops)
    img = cases.lod_batch(B, H, W, seed=31)
    rows = []
    for b, op in enumerate(ops):
        if op < 0:
            rows.append(torch.zeros((1, 24)))
        else:
            rows.append(AF.pack_params(cases.params_for(op, 1, seed=40 + b)[1], O.OP_NPARAMS[op]))
    P = torch.cat(rows, 0).to(dev)
    ops_t = torch.tensor(ops, dtype=torch.int32, device=dev)
    y, _, down = AF.apply_ops(img.to(dev), P, ops_t, clip=True, down_hw=out)
    y_plain = AF.apply_ops(img.to(dev), P, ops_t, clip=True)
    assert torch.equal(y, y_plain)                                   # the emitting kernels compute the same image
    ref = torch.nn.AdaptiveAvgPool2d(out)(y.double()).float()
    assert down.shape == ref.shape
    err = (down - ref).abs().amax(dim=(1, 2, 3)).cpu().tolist()
    assert max(err) <= 3e-7, err
    assert float(down[3].abs().max()) == 0.0                         # AISP_OP_NONE: zero image, zero means


@pytest.mark.parametrize("H,W,out", [(128, 128, (16, 16)), (96, 160, (32, 32)), (64, 256, (16, 128))])
def test_gradient_through_emitted_block_means(dev, H, W, out):
    """train.py:283: the critic pools the RETOUCHED image (value.py:63), so a gradient reaches the block
    means; it must act on the filter parameters exactly as if the image had been pooled by PyTorch.
    (128 -> 16: power-of-two blocks, the pooled gradient is added inside the backward kernels' loads;
    96x160 -> 32x32: 3x5 blocks, the up-sampled gradient is added up front; 64x256 -> 16x128: 4x2 blocks.)"""
    from adaptiveisp_b200 import functional as AF
    ops = [O.OP_EXPOSURE, O.OP_TONE, O.OP_SHARPEN, O.OP_NLM, O.OP_CCM, O.OP_USM, O.OP_SATPLUS]
    B = len(ops)
    img = cases.lod_batch(B, H, W, seed=51, device=dev)
    P0 = torch.cat([AF.pack_params(cases.params_for(op, 1, seed=60 + b)[1], O.OP_NPARAMS[op]) for b, op in enumerate(ops)], 0)
    ops_t = torch.tensor(ops, dtype=torch.int32, device=dev)
    g = cases.grad_out((B, 3, H, W), 5).to(dev)
    gd = cases.grad_out((B, 3) + tuple(out), 6).to(dev).abs() * 40.0      # comparable in size to the direct gradient
    Pa = P0.to(dev).requires_grad_(True)
    y, _, down = AF.apply_ops(img, Pa, ops_t, clip=True, down_hw=out)
    ((y * g).sum() + (down * gd).sum()).backward()
    Pb = P0.to(dev).requires_grad_(True)
    y2 = AF.apply_ops(img, Pb, ops_t, clip=True)
    ((y2 * g).sum() + (torch.nn.AdaptiveAvgPool2d(out)(y2) * gd).sum()).backward()
    for b, op in enumerate(ops):
        n = O.OP_NPARAMS[op]
        assert_grad(Pa.grad[b, :n].cpu().numpy(), Pb.grad[b, :n].cpu().numpy(), 2e-5, O.OP_NAMES[op])
    # only the pooled image is used: the full-resolution gradient is absent altogether
    Pc = P0.to(dev).requires_grad_(True)
    _, _, d3 = AF.apply_ops(img, Pc, ops_t, clip=True, down_hw=out)
    (d3 * gd).sum().backward()
    Pd = P0.to(dev).requires_grad_(True)
    (torch.nn.AdaptiveAvgPool2d(out)(AF.apply_ops(img, Pd, ops_t, clip=True)) * gd).sum().backward()
    for b, op in enumerate(ops):
        n = O.OP_NPARAMS[op]
        assert_grad(Pc.grad[b, :n].cpu().numpy(), Pd.grad[b, :n].cpu().numpy(), 5e-5, "pooled only " + O.OP_NAMES[op])


def test_agent_output_carries_block_means_for_the_critic(dev):
    """Agent.forward attaches the emitted block means to the retouched image; the next Agent step and the
    drop-in critic use them instead of pooling the full-resolution image again."""
    from adaptiveisp_b200.agent import Agent
    from adaptiveisp_b200.config import make_cfg
    from adaptiveisp_b200.value import Value
    torch.manual_seed(11)
    cfg = make_cfg(feature_extractor_dims=64, base_channels=4, fc1_size=16, dropout_keep_prob=1.0)
    agent = Agent(cfg, shape=(16, 64, 64), device=dev).to(dev).eval()
    value = Value(cfg, shape=(19, 64, 64)).to(dev).eval()
    B = 6
    x = cases.lod_batch(B, 512, 512, seed=71, device=dev)
    z = torch.rand((B, cfg.z_dim), device=dev)
    s0 = torch.zeros((B, cfg.num_state_dim), device=dev)
    with torch.no_grad():
        (y, s1, _sur, _pen), dbg, _ = agent((x, z, s0), 1.0)
        ref_down = torch.nn.AdaptiveAvgPool2d((64, 64))(y.double()).float()
        assert float((y._aisp_down - ref_down).abs().max()) <= 3e-7
        assert agent.downsample(y) is y._aisp_down
        v_fast = value(y, s1)
        v_ref = value(y.clone(), s1)                 # a plain tensor: pooled by the block-mean kernel
        assert float((v_fast - v_ref).abs().max()) <= 1e-5
        (y2, s2, _, _), _, _ = agent((y, z, s1), 1.0)   # second step consumes the emitted means
        (y2r, s2r, _, _), _, _ = agent((y.clone(), z, s1), 1.0)
        assert torch.equal(s2, s2r) and float((y2 - y2r).abs().max()) <= 1e-5


# ----------------------------------------------------------------------------------------------
# 8. differentiable replay: backward through a planned batch of pipelines
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("clip_each", [True, False])
def test_backward_through_planned_pipelines(dev, clip_each):
    """replay_grad.apply_plan: gradients w.r.t. the input image and w.r.t. EVERY step's parameters of
    heterogeneous pipelines (per-pixel runs fused, one stencil step per phase, two-stencil pipelines
    split into phases) against autograd through the CPU oracle; values equal execute_plan's."""
    from adaptiveisp_b200 import replay
    from adaptiveisp_b200.replay_grad import apply_plan
    seqs = [
        [O.OP_EXPOSURE, O.OP_GAMMA, O.OP_WB, O.OP_CCM, O.OP_SHARPEN],          # isp/filters.py:753-815
        [O.OP_WNB, O.OP_NLM, O.OP_USM],                                        # BASELINE configs[3]: two phases
        [O.OP_TONE, O.OP_USM, O.OP_CONTRAST, O.OP_SATPLUS],
        [O.OP_CCM, O.OP_COLOR],
        [O.OP_SHARPEN_V2],
    ]
    B, H, W = len(seqs), 24, 36
    img, _, _, _, plist = _seq_case(seqs, H, W, seed=77)
    plan = replay.plan_pipeline(seqs, plist, dev)
    assert len(plan.phases) == 2
    P = [ph.params.clone().requires_grad_(True) for ph in plan.phases]
    xd = img.to(dev).requires_grad_(True)
    g = cases.grad_out(img.shape, 77)
    g[1].abs_()                                       # the NLM sample: see test_filter_matches_oracle
    y = apply_plan(xd, plan, P, clip_each=clip_each)
    (y * g.to(dev)).sum().backward()
    fast = replay.execute_plan(img.to(dev), plan, clip_each=clip_each)
    assert float((fast - y.detach()).abs().max()) <= 1e-6
    segs = [replay.segment(s) for s in seqs]
    chk = Checks()
    for b, seq in enumerate(seqs):
        xc = img[b:b + 1].clone().requires_grad_(True)
        pcs = [p.clone().requires_grad_(True) for p in plist[b]]
        yc = O.chain(seq, xc, pcs, clip_each)
        (yc * g[b:b + 1]).sum().backward()
        tag = f"sample {b} {[O.OP_NAMES[o] for o in seq]}"
        tol_o, tol_g = 1e-5 * max(len(seq), 3), 1e-4 * max(len(seq), 3)
        chk.out(y[b:b + 1].detach().cpu().numpy(), yc.detach().numpy(), tol_o, tag)
        chk.norm(xd.grad[b:b + 1].cpu().numpy(), xc.grad.numpy(), tol_g, tag + " d/d img")
        for phase, seg in enumerate(segs[b]):
            for j, k in enumerate(seg):
                n = O.OP_NPARAMS[seq[k]]
                ref = pcs[k].grad.reshape(-1).numpy()
                if np.abs(ref).max() > 1e-6:
                    chk.grad(P[phase].grad[b, j, :n].cpu().numpy(), ref, tol_g, f"{tag} step {k}")
    chk.done()


# ----------------------------------------------------------------------------------------------
# 9. batched parameter prediction: one fc1 GEMM + one regressor kernel for the whole bank
# ----------------------------------------------------------------------------------------------
def test_batched_parameter_prediction_matches_the_modules(F, cfg, dev):
    """filters.BankPredictor (what Agent.forward and FilterBank use) against the per-module statement
    `extract_parameters` + `filter_param_regressor` of every drop-in class (all 13: cfg.filters plus
    USM, ColorFilter, SharpenFilterV2): parameters, and the gradients that reach each module's own
    fc1 / fc_filter weights; fc_mask gets none."""
    torch.manual_seed(7)
    classes = list(cfg.filters) + [F.SharpenUSMFilter, F.ColorFilter, F.SharpenFilterV2]
    mods = [c(cfg, predict=True).to(dev) for c in classes]
    B = 9
    feats = (torch.randn((B, cfg.feature_extractor_dims), device=dev) * 0.3).requires_grad_(True)
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        pred = F.BankPredictor(mods)
        P = pred(feats)
        gP = torch.randn(P.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
        (P * gP).sum().backward()
        got_w = [(m.fc1.weight.grad.clone(), m.fc_filter.weight.grad.clone(), m.fc_filter.bias.grad.clone()) for m in mods]
        got_f = feats.grad.clone()
        assert all(m.fc_mask.weight.grad is None for m in mods)
        for m in mods:
            m.zero_grad(set_to_none=True)
        feats.grad = None
        loss = 0.0
        for i, m in enumerate(mods):
            raw, _ = m.extract_parameters(feats)
            p = m.filter_param_regressor(raw)
            n = m.get_num_filter_parameters()
            flat = p.reshape(B, n)
            assert float((P[:, i, :n] - flat).abs().max()) <= 2e-6, m.get_short_name()
            assert float(P[:, i, n:].abs().max()) == 0.0 if n < 24 else True
            loss = loss + (flat * gP[:, i, :n]).sum()
        loss.backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    for m, (g1, gw, gb) in zip(mods, got_w):
        tag = m.get_short_name()
        assert rel_err(g1.cpu().numpy(), m.fc1.weight.grad.cpu().numpy()) <= 1e-4, tag
        assert rel_err(gw.cpu().numpy(), m.fc_filter.weight.grad.cpu().numpy()) <= 1e-4, tag
        assert rel_err(gb.cpu().numpy(), m.fc_filter.bias.grad.cpu().numpy()) <= 1e-4, tag
    assert rel_err(got_f.cpu().numpy(), feats.grad.cpu().numpy()) <= 1e-4
    # the reference layouts come back as views of the packed rows
    per = pred.split(P)
    assert per[classes.index(F.ToneFilter)].shape == (B, 8, 1, 1, 1) and per[classes.index(F.ColorFilter)].shape == (B, 8, 3, 1, 1)


def test_value_statistics_kernel_matches_the_pytorch_statement(dev):
    """value.py:64-75 on the pooled image: mean / unbiased variance of the luminance and the mean
    saturation, one launch, against the reference's own op sequence (value.value_statistics)."""
    from adaptiveisp_b200 import functional as AF
    from adaptiveisp_b200.value import value_statistics
    for (B, h, w) in [(5, 64, 64), (3, 16, 24), (2, 9, 11)]:
        x = cases.edge_image(B, h, w, seed=h).to(dev)          # samples >= 2 spill outside [0,1]: the clip matters
        got = AF.value_stats(x)
        ref = value_statistics(x.double()).float()
        ref32 = value_statistics(x)
        assert got.shape == (B, 3)
        err = ((got - ref).abs() / ref.abs().clamp_min(1e-3)).max()
        assert float(err) <= 1e-5, (B, h, w, got, ref)
        assert float(((got - ref32).abs() / ref32.abs().clamp_min(1e-3)).max()) <= 1e-4


# ----------------------------------------------------------------------------------------------
# 10. the bare NLM modules of isp/denoise.py (rows A11 / A12): NonLocalMeansGray and NonLocalMeans (RGB)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["gray", "rgb"])
@pytest.mark.parametrize("variant", ["a", "b"])
def test_bare_nlm_modules_match_reference_vectors(dev, name, variant):
    """adaptiveisp_b200.denoise.NonLocalMeansGray / NonLocalMeans (per-channel distances and weights) against
    vectors from the unmodified reference: images that spill outside [0,1] (the modules do not clip their
    input: the gray variant measures distances on the clipped luma but averages the raw image), outputs and
    the gradient w.r.t. h."""
    import os
    from adaptiveisp_b200 import denoise as D
    G = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "denoise_modules.npz")))
    key = f"{name}.{variant}"
    img, g = torch.from_numpy(G[key + ".img"]).to(dev), torch.from_numpy(G[key + ".g"]).to(dev)
    h = torch.from_numpy(G[key + ".h"]).to(dev).requires_grad_(True)
    mod = (D.NonLocalMeansGray if name == "gray" else D.NonLocalMeans)(search_window_size=11, patch_size=5)
    y = mod(img, h)
    (y * g).sum().backward()
    assert out_err(y.detach().cpu().numpy(), G[key + ".out"]) <= OUT_ATOL
    assert_grad(h.grad.cpu().numpy(), G[key + ".gh"], 2e-4, f"{name} d/dh")   # the reference's own fp32 noise is ~5e-5


@pytest.mark.parametrize("variant", ["a", "b"])
def test_nlm_param_module_matches_reference_vectors(dev, variant):
    """adaptiveisp_b200.denoise.NonLocalMeansParam (isp/denoise.py:122-157: reflect-padded search window, box
    as large as the window, one learnable scalar h) against vectors from the unmodified reference, and against
    the oracle on a larger image with the default-style window."""
    import os
    from adaptiveisp_b200 import denoise as D
    G = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "denoise_param.npz")))
    key = f"param.{variant}"
    img, g = torch.from_numpy(G[key + ".img"]).to(dev), torch.from_numpy(G[key + ".g"]).to(dev)
    S, h0 = int(G[key + ".cfg"][0]), float(G[key + ".cfg"][1])
    mod = D.NonLocalMeansParam(h0, search_window_size=S).to(dev)
    assert [k for k, _ in mod.named_parameters()] == ["h"] and mod.h.shape == (1,)
    y = mod(img)
    (y * g).sum().backward()
    assert out_err(y.detach().cpu().numpy(), G[key + ".out"]) <= OUT_ATOL
    assert_grad(mod.h.grad.cpu().numpy(), G[key + ".gh"], 2e-4, "param d/dh")
    if variant == "a":   # a larger, ragged image against the oracle
        x = cases.edge_image(2, 37, 45, seed=52)
        hc = torch.tensor([0.3], requires_grad=True)
        yc = O.nlm_param_module(x, hc, 9)
        yc.sum().backward()
        m2 = D.NonLocalMeansParam(0.3, search_window_size=9).to(dev)
        y2 = m2(x.to(dev))
        y2.sum().backward()
        assert out_err(y2.detach().cpu().numpy(), yc.detach().numpy()) <= OUT_ATOL
        assert_grad(m2.h.grad.cpu().numpy(), hc.grad.numpy(), 2e-4, "param d/dh (oracle)")


def test_bare_nlm_modules_api(dev):
    from adaptiveisp_b200 import AispError, denoise as D
    with pytest.raises(AispError):
        D.NonLocalMeansGray(search_window_size=21, patch_size=7)       # only the 11 / 5 configuration is built
    with pytest.raises(AispError):
        D.NonLocalMeansParam(0.5, search_window_size=4)                # windows are odd
    with pytest.raises(AispError):
        D.NonLocalMeansParam(0.5, search_window_size=21).to(dev)(cases.lod_batch(1, 8, 64, seed=1, device=dev))   # pad >= H
    x = cases.lod_batch(2, 48, 64, seed=3, device=dev)
    # scalar h shared by the batch: the gradient is summed over the samples
    h = torch.tensor([0.3], device=dev, requires_grad=True)
    y = D.NonLocalMeans(11, 5)(x, h)
    y.sum().backward()
    hb = torch.tensor([0.3, 0.3], device=dev).reshape(2, 1, 1, 1).requires_grad_(True)
    yb = D.NonLocalMeans(11, 5)(x, hb)
    yb.sum().backward()
    assert torch.equal(y, yb) and abs(float(h.grad) - float(hb.grad.sum())) <= 1e-4 * abs(float(h.grad)) + 1e-6
    with pytest.raises(AispError):
        D.NonLocalMeansGray(11, 5)(x.clone().requires_grad_(True), h)
    # the small helpers are the reference's torch statements
    lum = D.rgb_to_luminance(x)
    assert lum.shape == (2, 1, 48, 64)
    assert D.BoxFilter(5, "sum")(lum).shape == lum.shape and D.ShiftStack(3)(lum).shape == (2, 1, 48, 64, 9)


# ----------------------------------------------------------------------------------------------
# 11. shot / read noise of the synthetic-RAW model (isp/unprocess_np.py:131-181) on the resident batch
# ----------------------------------------------------------------------------------------------
def test_shot_read_noise_matches_reference_vectors(dev):
    """aisp_shot_read_noise with the reference's own normals: brightness ratio, then x + sqrt(x*shot + read)*z
    against the unmodified NumPy module's output (float64 there, fp32 here: 1e-6)."""
    import os
    from adaptiveisp_b200 import unprocess as U
    G = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "noise_model.npz")))
    img = torch.from_numpy(G["img"]).to(dev)
    z = torch.from_numpy(G["z"].astype(np.float32)).to(dev)
    y = U.add_read_and_shot_noise(img, G["shot"], G["read"], gain=G["gain"], z=z)
    assert np.abs(y.cpu().numpy().astype(np.float64) - G["out"]).max() <= 1e-6
    # ragged / unaligned sizes take the scalar path; in place is allowed
    x = torch.rand((2, 3, 7, 9), device=dev)
    zz = torch.randn_like(x)
    ref = O.shot_read_noise(x.cpu().numpy(), [0.003, 0.01], [1e-4, 2e-5], zz.cpu().numpy())
    y2 = U.add_read_and_shot_noise(x.clone(), [0.003, 0.01], [1e-4, 2e-5], z=zz)
    assert np.abs(y2.cpu().numpy() - ref).max() <= 1e-6
    xi = x.clone()
    U.add_read_and_shot_noise(xi, [0.003, 0.01], [1e-4, 2e-5], z=zz, out=xi)
    assert torch.equal(xi, y2)
    # a negative variance is NaN, as numpy's sqrt gives
    neg = torch.full((1, 3, 4, 4), -1.0, device=dev)
    assert torch.isnan(U.add_read_and_shot_noise(neg, 0.5, 0.1, z=torch.ones_like(neg))).all()


def test_shot_read_noise_in_kernel_normals(dev):
    """Philox normals generated in the kernel: deterministic in (seed, offset), standard-normal moments, and the
    heteroscedastic variance x * shot + read of isp/unprocess_np.py:178-181 per image."""
    from adaptiveisp_b200 import unprocess as U
    B, n = 3, 1 << 20
    level = torch.tensor([0.05, 0.3, 0.8], device=dev).reshape(B, 1).expand(B, n).contiguous()
    shot, read = [0.01, 0.002, 0.012], [1e-4, 5e-6, 3e-4]
    a = U.add_read_and_shot_noise(level, shot, read, seed=7)
    b = U.add_read_and_shot_noise(level, shot, read, seed=7)
    c = U.add_read_and_shot_noise(level, shot, read, seed=7, offset=n // 4)
    d = U.add_read_and_shot_noise(level, shot, read, seed=8)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, d)
    for i in range(B):
        var = float(level[i, 0]) * shot[i] + read[i]
        zhat = ((a[i] - level[i]) / var ** 0.5).double()
        assert abs(float(zhat.mean())) <= 5e-3                       # 5 sigma of the mean of 2^20 normals
        assert abs(float(zhat.var()) - 1.0) <= 1e-2
        assert abs(float((zhat ** 4).mean()) - 3.0) <= 5e-2          # kurtosis of a normal
        assert abs(float((zhat[:-1] * zhat[1:]).mean())) <= 5e-3     # neighbours uncorrelated
    # images get different streams
    assert abs(float(((a[0] - level[0]) * (a[1] - level[1])).mean())) <= 1e-5
