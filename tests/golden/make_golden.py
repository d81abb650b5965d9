"""Generate tests/golden/*.npz by running the UNMODIFIED reference (read-only checkout, build
container only) on the seeded edge-case inputs of tests/cases.py.

    python tests/golden/make_golden.py          # rewrites tests/golden/filters.npz, select.npz

Stored per filter class: the input image, the raw FC features, the regressed parameters
(``filter_param_regressor``), ``process()`` output (== ``run()``), the clipped ``forward()`` output, and
the autograd gradients of sum(g * out) w.r.t. parameters and image for both variants.
For the selection logic: pdf / noise / states -> pdf_sample ids, one_hot, and Agent.forward's
new_states for a seeded Agent (states only; the nets' weights are not stored).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import ref_shim, cases  # noqa: E402
from oracle import isp_oracle as O  # noqa: E402

torch.set_num_threads(1)  # fixed reduction order for the stored reference values


def ref_class(ref, op):
    F = ref.filters
    return {
        O.OP_EXPOSURE: F.ExposureFilter, O.OP_GAMMA: F.GammaFilter, O.OP_CCM: F.CCMFilter,
        O.OP_SHARPEN: F.SharpenFilter, O.OP_NLM: F.DenoiseFilter, O.OP_TONE: F.ToneFilter,
        O.OP_CONTRAST: F.ContrastFilter, O.OP_SATPLUS: F.SaturationPlusFilter, O.OP_WNB: F.WNBFilter,
        O.OP_WB: F.ImprovedWhiteBalanceFilter, O.OP_USM: F.SharpenUSMFilter, O.OP_COLOR: F.ColorFilter,
        O.OP_SHARPEN_V2: F.SharpenFilterV2,
    }[op]


def filters_golden(ref):
    out = {}
    for op in cases.ALL_OPS:
        name = O.OP_NAMES[op]
        flt = ref_class(ref, op)(ref.cfg, predict=False)
        for variant, (B, H, W, seed) in {"a": (3, 20, 24, 0), "b": (1, 13, 17, 1)}.items():
            img = cases.edge_image(B, H, W, seed)
            feat, _ = cases.params_for(op, B, seed)
            g = cases.grad_out(img.shape, seed)
            key = f"{name}.{variant}"
            out[key + ".img"] = img.numpy()
            out[key + ".feat"] = feat.numpy()
            out[key + ".g"] = g.numpy()
            # regressor as the reference computes it, from the same (possibly biased) features
            param = flt.filter_param_regressor(feat.clone())
            out[key + ".param"] = param.detach().numpy()
            for mode in ("run", "fwd"):
                x = img.clone().requires_grad_(True)
                p = param.detach().clone().requires_grad_(True)
                if mode == "run":
                    y = flt.run(x, p)
                else:
                    y, _, _ = flt.forward(x, specified_parameter=p)
                (y * g).sum().backward()
                out[f"{key}.{mode}.out"] = y.detach().numpy()
                out[f"{key}.{mode}.gparam"] = p.grad.numpy()
                out[f"{key}.{mode}.gimg"] = x.grad.numpy()
    # known-answer vector from SURVEY.md §8a: gray pixel through S+ gets tinted
    flt = ref.filters.SaturationPlusFilter(ref.cfg)
    kat = flt.process(torch.full((1, 3, 1, 1), 0.25), torch.tensor([[0.5]]))
    out["S+.kat.out"] = kat.numpy()
    return out


def select_golden(ref):
    out = {}
    A = ref.agent
    g = torch.Generator().manual_seed(77)
    B, n = 64, 10
    logits = torch.randn((B, n), generator=g) * 2.0
    u = torch.rand((B, 1), generator=g)
    u[0, 0] = 0.0
    u[1, 0] = 1.0
    pdf = torch.softmax(logits, dim=1) + 1e-37
    pdf = pdf * (1 - 0.05) + 0.05 * 1.0 / n
    pdf = pdf / (torch.sum(pdf, dim=1, keepdim=True) + 1e-30)
    ids = A.pdf_sample(pdf, u)
    out["logits"], out["u"], out["pdf"] = logits.numpy(), u.numpy(), pdf.numpy()
    out["ids_sample"] = ids.numpy()
    out["ids_argmax"] = torch.argmax(pdf, dim=1).to(torch.int32).numpy()
    out["one_hot"] = A.one_hot(n, ids.to(torch.int64)).numpy()

    # a live Agent.forward on tiny frames: states -> new_states / ids bookkeeping
    torch.manual_seed(5)
    agent = A.Agent(ref.cfg, shape=(16, 64, 64), device="cpu")
    Bs = 6
    x = cases.edge_image(Bs, 64, 64, seed=3, in_range=True)
    z = torch.rand((Bs, ref.cfg.z_dim), generator=g)
    states = torch.zeros((Bs, ref.cfg.num_state_dim))
    states[:, 2] = torch.tensor([0, 1, 2, 3, 4, 4.0])
    states[1, 3 + 2] = 1.0
    states[3, 3:] = 1.0
    for mode in ("train", "eval"):
        agent.train(mode == "train")
        torch.manual_seed(9)
        with torch.no_grad():
            (xo, new_states, surrogate, penalty), dbg, _ = agent((x, z, states), 0.3)
        sel = dbg["selected_filter"]
        out[f"agent.{mode}.x"] = x.numpy()
        out[f"agent.{mode}.u"] = z[:, 0:1].numpy()
        out[f"agent.{mode}.states"] = states.numpy()
        out[f"agent.{mode}.selected"] = sel.numpy()
        out[f"agent.{mode}.new_states"] = new_states.numpy()
    return out


def small_agent_cfg(cfg_cls, base):
    """A narrow Agent (about 25k parameters) so that its state_dict fits in a fixture."""
    small = cfg_cls(base)
    small.feature_extractor_dims = 64
    small.base_channels = 4
    small.fc1_size = 16
    small.dropout_keep_prob = 1.0  # no dropout: keeps train mode free of device-specific RNG
    return small


def agent_golden(ref):
    """Reference Agent.forward (+ backward) on a narrow Agent: weights, inputs, outputs, gradients."""
    out = {}
    A = ref.agent
    cfg = small_agent_cfg(type(ref.cfg), ref.cfg)
    torch.manual_seed(11)
    agent = A.Agent(cfg, shape=(16, 64, 64), device="cpu")
    # give every BatchNorm non-trivial running stats / affine so that eval mode is not degenerate
    g = torch.Generator().manual_seed(12)
    for m in agent.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    for k, v in agent.state_dict().items():
        out["sd." + k] = v.numpy().copy()   # copy: .numpy() aliases the live buffers, which train-mode forwards mutate
    B = 4
    x = cases.edge_image(B, 32, 32, seed=7, in_range=True)
    z = torch.rand((B, cfg.z_dim), generator=g)
    states = torch.zeros((B, cfg.num_state_dim))
    states[:, 2] = torch.tensor([0.0, 2.0, 4.0, 1.0])
    states[1, 3 + 4] = 1.0
    gout = cases.grad_out(x.shape, seed=7)
    out["x"], out["z"], out["states"], out["gout"] = x.numpy(), z.numpy(), states.numpy(), gout.numpy()

    sd0 = {k: v.clone() for k, v in agent.state_dict().items()}

    def run(tag, train, forced):
        agent.load_state_dict(sd0)      # train-mode forwards move the BatchNorm running statistics
        agent.train(train)
        agent.zero_grad(set_to_none=True)
        (xo, ns, sur, pen), dbg, _ = agent((x, z, states), 0.25, None, forced)
        loss = (xo * gout).sum() + sur.sum() + pen.sum()
        loss.backward()
        out[f"{tag}.x"] = xo.detach().numpy()
        out[f"{tag}.new_states"] = ns.detach().numpy()
        out[f"{tag}.surrogate"] = sur.detach().numpy()
        out[f"{tag}.penalty"] = pen.detach().numpy()
        out[f"{tag}.selected"] = dbg["selected_filter"].numpy()
        out[f"{tag}.pdf0"] = dbg["pdf"].detach().numpy()
        for name, prm in agent.named_parameters():
            if prm.grad is not None:
                out[f"{tag}.grad.{name}"] = prm.grad.numpy().copy()
            else:
                out[f"{tag}.nograd.{name}"] = np.zeros(1)

    run("train", True, None)
    run("eval", False, None)
    for f in range(len(cfg.filters)):
        run(f"forced{f}", False, f)

    # 5-step rollouts as in yolov3/val_adaptiveisp.py:288-309 (eval mode: argmax selection) and with
    # sampled selection (train mode, as the replay-memory loop of train.py:234-381 carries images on)
    for tag, train in (("rollout_eval", False), ("rollout_train", True)):
        agent.load_state_dict(sd0)
        agent.train(train)
        noises = torch.rand((cfg.test_steps, B, cfg.z_dim), generator=g)
        st = torch.zeros((B, cfg.num_state_dim))
        cur = x
        out[f"{tag}.noises"] = noises.numpy()
        with torch.no_grad():
            for i in range(cfg.test_steps):
                (cur, st, _, _), dbg, _ = agent((cur, noises[i], st), 1.0, None, None)
                out[f"{tag}.sel{i}"] = dbg["selected_filter"].numpy()
                out[f"{tag}.x{i}"] = cur.numpy()
                out[f"{tag}.states{i}"] = st.numpy()
                if st[0][1] > 0:
                    break
        out[f"{tag}.steps_run"] = np.array([i + 1])
    return out


def denoise_modules_golden(ref):
    """The bare NLM modules of isp/denoise.py: NonLocalMeansGray and NonLocalMeans
    (per-channel distances) with search 11 / patch 5, on images that spill outside [0,1] (the modules do not
    clip their input), with the gradient w.r.t. h under a one-signed upstream gradient."""
    out = {}
    for name, cls in (("gray", ref.denoise.NonLocalMeansGray), ("rgb", ref.denoise.NonLocalMeans)):
        for variant, (B, H, W, seed) in {"a": (3, 20, 24, 40), "b": (2, 33, 47, 41)}.items():
            img = cases.edge_image(B, H, W, seed)                    # samples >= 2 spill outside [0,1]
            h = (0.15 + 0.5 * torch.rand((B, 1, 1, 1), generator=torch.Generator().manual_seed(seed))).requires_grad_(True)
            g = cases.grad_out(img.shape, seed).abs()
            y = cls(search_window_size=11, patch_size=5)(img, h)
            (y * g).sum().backward()
            key = f"{name}.{variant}"
            out[key + ".img"], out[key + ".h"], out[key + ".g"] = img.numpy(), h.detach().numpy(), g.numpy()
            out[key + ".out"], out[key + ".gh"] = y.detach().numpy(), h.grad.numpy()
    return out


def denoise_param_golden(ref):
    """NonLocalMeansParam (isp/denoise.py:122-157; instantiated nowhere in the reference): reflect-padded
    windows of 5 and 7 on small images that spill outside [0,1], output and the gradient of the scalar h."""
    out = {}
    for variant, (B, H, W, S, h0, seed) in {"a": (2, 14, 17, 5, 0.35, 50), "b": (1, 19, 16, 7, 0.2, 51)}.items():
        img = cases.edge_image(B, H, W, seed)
        g = cases.grad_out(img.shape, seed).abs()
        mod = ref.denoise.NonLocalMeansParam(h0, search_window_size=S)
        y = mod(img)
        (y * g).sum().backward()
        key = f"param.{variant}"
        out[key + ".img"], out[key + ".g"] = img.numpy(), g.numpy()
        out[key + ".cfg"] = np.array([S, h0], dtype=np.float64)
        out[key + ".out"], out[key + ".gh"] = y.detach().numpy(), mod.h.grad.numpy()
    return out


def noise_model_golden():
    """Noise model of isp/unprocess_np.py:131-181 from the unmodified module (NumPy RNG, seeded): per-image
    levels of random_noise_levels_log, brightness ratios of adjust_random_brightness, and
    add_read_and_shot_noise on the scaled frames together with the standard normals it drew."""
    import importlib.util
    import sys
    import types
    for name in ("matplotlib", "matplotlib.pyplot"):      # imported at module level for a debug viewer only
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    spec = importlib.util.spec_from_file_location("ref_unprocess", os.path.join(ref_shim.REF_ROOT, "isp", "unprocess_np.py"))
    ru = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ru)
    np.random.seed(21)
    img = np.random.rand(3, 3, 10, 12).astype(np.float32)
    np.random.seed(22)
    levels = [ru.random_noise_levels_log() for _ in range(3)]
    np.random.seed(23)
    ratio = [ru.adjust_random_brightness(img[b], (0.1, 0.3))[1] for b in range(3)]
    np.random.seed(24)
    out = np.stack([ru.add_read_and_shot_noise(img[b].astype(np.float64) * ratio[b], levels[b][0], levels[b][1])
                    for b in range(3)])
    np.random.seed(24)
    z = np.random.normal(0, 1, img.shape)                  # the normals np.random.normal(0, scale) scaled
    np.random.seed(25)
    lin = ru.random_noise_levels_linear()
    return dict(img=img, shot=np.array([l[0] for l in levels]), read=np.array([l[1] for l in levels]),
                gain=np.array(ratio), z=z, out=out, linear=np.array(lin))


def main():
    ref = ref_shim.load()
    a = agent_golden(ref)
    np.savez_compressed(os.path.join(HERE, "agent.npz"), **a)
    print("agent.npz", os.path.getsize(os.path.join(HERE, "agent.npz")), "bytes")
    f = filters_golden(ref)
    np.savez_compressed(os.path.join(HERE, "filters.npz"), **f)
    s = select_golden(ref)
    np.savez_compressed(os.path.join(HERE, "select.npz"), **s)
    d = denoise_modules_golden(ref)
    np.savez_compressed(os.path.join(HERE, "denoise_modules.npz"), **d)
    np.savez_compressed(os.path.join(HERE, "denoise_param.npz"), **denoise_param_golden(ref))
    np.savez_compressed(os.path.join(HERE, "noise_model.npz"), **noise_model_golden())
    for fn in ("filters.npz", "select.npz"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")


if __name__ == "__main__":
    main()
