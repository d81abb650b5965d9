"""CPU-only checks: the C-ABI library builds/loads and exports every symbol the header declares,
host-side packing / rejection logic, drop-in class surface and checkpoint-key compatibility.
No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import isp_oracle as O
from tests import cases, ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from adaptiveisp_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "aisp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aisp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from adaptiveisp_b200 import _lib
    syms = header_symbols()
    assert len(syms) == 27
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/aisp_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes binding and header disagree"


def test_library_metadata_calls(lib):
    assert lib.aisp_version() == 3
    for op in range(13):
        assert lib.aisp_op_num_params(op) == O.OP_NPARAMS[op]
    assert lib.aisp_op_num_params(99) == -1
    assert lib.aisp_bwd_scratch_bytes(64, 512, 512) == 64 * 128 * 32 * 4 * 6    # x AISP_MAX_CHAIN_BWD
    assert lib.aisp_bwd_scratch_bytes(0, 512, 512) == 0
    assert b"NULL" in lib.aisp_status_string(-1)
    # argument validation happens before any CUDA call, so it is checkable without a GPU
    assert lib.aisp_pointwise_fwd(None, None, None, None, None, 1, 8, 8, 1, 1, None) == -1
    buf = (ctypes.c_float * 16)()
    p = ctypes.addressof(buf)
    assert lib.aisp_pointwise_fwd(p, p + 4, p, p, None, 0, 8, 8, 1, 1, None) == -2
    assert lib.aisp_pointwise_fwd(p, p + 4, p, p, None, 1, 8, 8, 9, 1, None) == -2
    assert lib.aisp_pointwise_fwd(p, p, p, p, None, 1, 2, 2, 1, 1, None) == -4       # in-place refused
    assert lib.aisp_pointwise_bwd(p, p, p, p, 1, 2, 2, 1, p, None, p, 4, None) == -3  # scratch too small
    assert lib.aisp_nlm_bwd(p, p, p, 1, 2, 2, p, p, 4, None) == -3                     # scratch too small
    assert lib.aisp_nlm_bwd_img(p, p, None, p, p, p, 1, 2, 2, p, None) == -1          # wsum stash is required


def test_op_codes_shared_between_header_oracle_and_host():
    from adaptiveisp_b200 import functional as AF
    text = open(os.path.join(ROOT, "include", "aisp_b200.h")).read()
    enum = dict((k, int(v)) for k, v in re.findall(r"AISP_OP_([A-Z0-9_]+)\s*=\s*(-?\d+)", text))
    for name, code in [("EXPOSURE", O.OP_EXPOSURE), ("GAMMA", O.OP_GAMMA), ("CCM", O.OP_CCM),
                       ("SHARPEN", O.OP_SHARPEN), ("NLM", O.OP_NLM), ("TONE", O.OP_TONE),
                       ("CONTRAST", O.OP_CONTRAST), ("SATPLUS", O.OP_SATPLUS), ("WNB", O.OP_WNB), ("WB", O.OP_WB),
                       ("USM", O.OP_USM), ("COLOR", O.OP_COLOR), ("SHARPEN_V2", O.OP_SHARPEN_V2)]:
        assert enum[name] == code == getattr(AF, "OP_" + name)
    assert tuple(O.OP_NPARAMS[i] for i in range(13)) == AF.NUM_PARAMS


def test_pack_params_layouts():
    from adaptiveisp_b200 import functional as AF
    for op in cases.ALL_OPS:
        _, p = cases.params_for(op, 3)
        row = AF.pack_params(p, O.OP_NPARAMS[op])
        assert row.shape == (3, 24)
        np.testing.assert_array_equal(row[:, :O.OP_NPARAMS[op]].numpy(), cases.flat(op, p).numpy())
        assert float(row[:, O.OP_NPARAMS[op]:].abs().sum()) == 0.0


def test_bank_and_select_argument_validation(lib):
    """The bank's op list is a HOST array: everything about it is decided before the first launch."""
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    I32 = ctypes.c_int32
    ok3 = (I32 * 3)(O.OP_EXPOSURE, O.OP_NLM, O.OP_SHARPEN)
    assert lib.aisp_bank_fwd(None, p, p, ok3, 1, 3, 8, 8, 1, None, None) == -1
    assert lib.aisp_bank_fwd(p, p + 4, p, None, 1, 3, 8, 8, 1, None, None) == -1
    assert lib.aisp_bank_fwd(p, p, p, ok3, 1, 3, 8, 8, 1, None, None) == -4                    # in place
    assert lib.aisp_bank_fwd(p, p + 4, p, ok3, 1, 0, 8, 8, 1, None, None) == -2                # F < 1
    assert lib.aisp_bank_fwd(p, p + 4, p, (I32 * 17)(*[0] * 17), 1, 17, 8, 8, 1, None, None) == -2   # F > 16
    assert lib.aisp_bank_fwd(p, p + 4, p, ok3, 30000, 3, 8, 8, 1, None, None) == -2            # B*F > 65535
    two_nlm = (I32 * 2)(O.OP_NLM, O.OP_NLM)
    assert lib.aisp_bank_fwd(p, p + 4, p, two_nlm, 1, 2, 8, 8, 1, None, None) == -4            # one stash per image
    assert lib.aisp_bank_fwd(p, p + 4, p, (I32 * 1)(-1), 1, 1, 8, 8, 1, None, None) == -4      # NONE is not a slot
    assert lib.aisp_bank_fwd(p, p + 4, p, (I32 * 1)(13), 1, 1, 8, 8, 1, None, None) == -4
    assert lib.aisp_bank_bwd(p, p, p, ok3, 1, 3, 8, 8, 1, None, p, p, 4, None) == -3           # scratch for B*F samples
    assert lib.aisp_bwd_scratch_bytes(64 * 10, 512, 512) == 10 * lib.aisp_bwd_scratch_bytes(64, 512, 512)
    # aisp_select: S must be 3 + F, sampling needs the noise, forced ids must name a filter
    i64 = ctypes.addressof((ctypes.c_int64 * 64)())
    args = lambda noise, mode, forced, S: (p, noise, mode, forced, p, p, p, 2, 10, S, 5.0, 1.0, i64, i64, p, p, p, p, None)
    assert lib.aisp_select(*args(None, 0, 0, 13)) == -1
    assert lib.aisp_select(*args(p, 0, 0, 12)) == -2
    assert lib.aisp_select(*args(p, 3, 0, 13)) == -4
    assert lib.aisp_select(*args(None, 2, 10, 13)) == -2
    assert lib.aisp_select_bwd(None, i64, 2, 10, p, None) == -1


def test_filter_bank_host_side():
    """FilterBank is a view over already built modules: op list in cfg.filters order, parameters
    regressed per filter in the reference's layouts (no kernel call here)."""
    from adaptiveisp_b200 import filters
    from adaptiveisp_b200.config import make_cfg
    cfg = make_cfg()
    mods = [c(cfg, predict=True) for c in cfg.filters]
    bank = filters.FilterBank(mods)
    assert bank.ops == [O.OP_EXPOSURE, O.OP_GAMMA, O.OP_CCM, O.OP_SHARPEN, O.OP_NLM, O.OP_TONE, O.OP_CONTRAST,
                        O.OP_SATPLUS, O.OP_WNB, O.OP_WB]
    feats = torch.randn(3, cfg.feature_extractor_dims)
    params = bank.parameters_for(img_features=feats)
    for m, prm in zip(mods, params):
        assert prm.shape[0] == 3 and prm[0].numel() == m.get_num_filter_parameters()
        assert torch.equal(prm, m.filter_param_regressor(m.extract_parameters(feats)[0]))
    assert bank.parameters_for(specified_parameters=params) == params
    with pytest.raises(ValueError):
        filters.FilterBank([])
    with pytest.raises(Exception, match="no CPU path|CUDA"):
        bank(torch.zeros((3, 3, 8, 8)), img_features=feats)


def test_cpu_tensors_are_rejected_not_converted():
    from adaptiveisp_b200 import AispError, filters
    from adaptiveisp_b200.config import make_cfg
    flt = filters.GammaFilter(make_cfg())
    with pytest.raises(AispError, match="no CPU path"):
        flt.process(torch.zeros((1, 3, 4, 4)), torch.ones((1, 1)))


def test_dropin_class_surface_and_regressors():
    """Same class names / short names / parameter counts / regressor values as the reference."""
    from adaptiveisp_b200 import filters as F
    from adaptiveisp_b200.config import make_cfg
    cfg = make_cfg()
    expect = {"ExposureFilter": ("E", 1), "GammaFilter": ("G", 1), "ImprovedWhiteBalanceFilter": ("W", 3),
              "ColorFilter": ("C", 24), "ToneFilter": ("T", 8), "ToneFilterV2": ("T", 8), "ContrastFilter": ("Ct", 1),
              "WNBFilter": ("BW", 1), "SaturationPlusFilter": ("S+", 1), "DenoiseFilter": ("NLM", 1),
              "SharpenUSMFilter": ("USM", 2), "SharpenFilter": ("Shr", 1), "SharpenFilterV2": ("Shr", 1),
              "CCMFilter": ("CCM", 9)}
    for name, (short, n) in expect.items():
        f = getattr(F, name)(cfg, predict=True)
        assert f.get_short_name() == short and f.get_num_filter_parameters() == n
        assert f.get_num_mask_parameters() == 6 and f.channels == 3 and not f.use_masking()
        assert sorted(k for k, _ in f.named_parameters()) == [
            "fc1.bias", "fc1.weight", "fc_filter.bias", "fc_filter.weight", "fc_mask.bias", "fc_mask.weight"]
        feat = cases.features(f.OP, 5, seed=2)
        got = f.filter_param_regressor(feat)
        want = O.regress(f.OP, feat)
        assert got.shape == want.shape
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=1e-7)


def test_agent_state_dict_keys_match_reference_checkpoint_layout(tmp_path):
    from adaptiveisp_b200.agent import Agent
    from adaptiveisp_b200.config import make_cfg
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "agent.npz")))
    ref_keys = sorted(k[3:] for k in g if k.startswith("sd."))
    cfg = make_cfg(feature_extractor_dims=64, base_channels=4, fc1_size=16, dropout_keep_prob=1.0)
    agent = Agent(cfg, shape=(16, 64, 64), device="cpu")
    assert sorted(agent.state_dict().keys()) == ref_keys
    for k, v in agent.state_dict().items():
        assert tuple(v.shape) == tuple(g["sd." + k].shape), k


def test_agent_selection_logic_bit_exact_on_cpu(golden_select):
    from adaptiveisp_b200 import agent as A
    G = golden_select
    pdf, u = torch.from_numpy(G["pdf"]), torch.from_numpy(G["u"])
    ids = A.pdf_sample(pdf, u)
    np.testing.assert_array_equal(ids.numpy(), G["ids_sample"])
    np.testing.assert_array_equal(A.one_hot(10, ids.to(torch.int64)).numpy(), G["one_hot"])
    assert int(ids.min()) == -1  # u == 0 row: all-zero one-hot, as in the reference
    assert int(A.one_hot(10, ids.to(torch.int64))[int(ids.argmin())].sum()) == 0


def test_lod_batch_shape_and_letterbox():
    x = cases.lod_batch(3, 48, 48, seed=1)
    assert x.shape == (3, 3, 48, 48) and x.dtype == torch.float32
    assert float(x[:, :, :8].abs().max()) == 0.0 and float(x[:, :, 40:].abs().max()) == 0.0
    assert float(x.max()) <= 1.0 and float(x.min()) >= 0.0
    np.testing.assert_allclose((x * 255).round().numpy(), (x * 255).numpy(), atol=1e-4)  # k/255 lattice


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")
def test_reference_config_accepts_dropin_classes():
    """config.py's star-import seam: the reference cfg object drives the drop-in classes unchanged."""
    from adaptiveisp_b200 import filters as F
    ref = ref_shim.load()
    for cls in ref.cfg.filters:
        mine = getattr(F, cls.__name__)(ref.cfg, predict=True)
        theirs = cls(ref.cfg, predict=True)
        assert mine.get_short_name() == theirs.get_short_name()
        assert {k: tuple(v.shape) for k, v in mine.state_dict().items()} == \
               {k: tuple(v.shape) for k, v in theirs.state_dict().items()}


def test_replay_segmentation_and_reference_file_formats():
    from adaptiveisp_b200 import replay
    from adaptiveisp_b200.config import make_cfg
    E, G, CCM, SHR, NLM, T = O.OP_EXPOSURE, O.OP_GAMMA, O.OP_CCM, O.OP_SHARPEN, O.OP_NLM, O.OP_TONE
    assert replay.segment([E, G, CCM]) == [[0, 1, 2]]
    # a segment = per-pixel steps around AT MOST ONE stencil step (one aisp_sequence_fwd pass over HBM)
    assert replay.segment([E, SHR, G, T, NLM]) == [[0, 1, 2, 3], [4]]
    assert replay.segment([E, G, SHR]) == [[0, 1, 2]]                       # isp/filters.py:753-815 style chain
    assert replay.segment([NLM, NLM]) == [[0], [1]]
    assert replay.segment([]) == []
    assert [len(s) for s in replay.segment([E] * 11)] == [8, 3]            # AISP_MAX_STEPS per fused pass
    # param_results/<img>.json as written by yolov3/val_adaptiveisp.py:301-327
    cfg = make_cfg()
    text = '{"pipeline": [2, 5, 0], "CCM": [1.5, -0.2, -0.3, 0.1, 1.2, -0.3, 0.0, -0.4, 1.4], ' \
           '"T": [[[[0.6]]], [[[0.7]]], [[[0.8]]], [[[0.9]]], [[[1.0]]], [[[1.1]]], [[[1.2]]], [[[1.3]]]], "E": [0.25]}'
    ops, plist = replay.parse_param_results(text, cfg.filters)
    assert ops == [O.OP_CCM, O.OP_TONE, O.OP_EXPOSURE]
    assert [p.numel() for p in plist] == [9, 8, 1] and float(plist[1][3]) == pytest.approx(0.9)
    header, rec = replay.parse_records("E,G,CCM,Shr,NLM,T,Ct,S+,BW,W\nimg1.png,2,5,0,-1,-1\nimg2.png,4,4,4,4,4\n")
    assert header[4] == "NLM" and rec == {"img1.png": [2, 5, 0], "img2.png": [4, 4, 4, 4, 4]}


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path through the oracle port) must run without
    a GPU and print ONE JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "isp_chain_megapixels_per_sec_fwd_bwd" and d["unit"] == "MP/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_value_dropin_matches_reference_on_cpu():
    """The critic stays PyTorch; the drop-in must load the reference's state_dict strictly and, on CPU
    tensors (no kernels involved), reproduce its output bit for bit -- pooled image, luminance /
    contrast / saturation statistics (value.py:63-75) and the nets."""
    if not ref_shim.available():
        pytest.skip("reference checkout not present")
    from adaptiveisp_b200.config import make_cfg
    from adaptiveisp_b200.value import Value, value_statistics
    ref = ref_shim.load()
    cfg = make_cfg(feature_extractor_dims=64, base_channels=4, fc1_size=16)
    torch.manual_seed(0)
    theirs = ref.value.Value(cfg, shape=(19, 64, 64)).eval()
    ours = Value(cfg, shape=(19, 64, 64)).eval()
    ours.load_state_dict(theirs.state_dict(), strict=True)
    x = cases.lod_batch(3, 128, 192, seed=3)
    states = torch.rand((3, cfg.num_state_dim))
    with torch.no_grad():
        assert torch.equal(ours(x, states), theirs(x, states))
    # an image that carries emitted block means uses them instead of pooling again
    x2 = x.clone()
    x2._aisp_down = torch.nn.AdaptiveAvgPool2d((64, 64))(x)
    with torch.no_grad():
        assert torch.equal(ours(x2, states), theirs(x, states))
    st = value_statistics(torch.nn.AdaptiveAvgPool2d((64, 64))(x))
    assert st.shape == (3, 3) and bool((st[:, 1] >= 0).all())


def test_replay_pool_reports_exhaustion_instead_of_spinning():
    """get_batch with every record checked out (no put_back / discard in between) must raise, not loop."""
    import random
    from adaptiveisp_b200.config import make_cfg
    from adaptiveisp_b200.replay_pool import DeviceReplayPool
    cfg = make_cfg(replay_memory_size=6)
    n = {"k": 0}

    def fetch(k):
        n["k"] += k
        return torch.zeros((k, 3, 4, 4)), [{"id": i} for i in range(k)]

    pool = DeviceReplayPool(cfg, (3, 4, 4), "cpu", fetch, fetch_batch=3, rng=random.Random(0))
    a = pool.get_batch(4)
    with pytest.raises(RuntimeError, match="exhausted"):
        pool.get_batch(4)                       # only 2 records left and nothing to refill
    assert len(pool) == 2                       # the failed draw put its picks back
    pool.put_back(a.slots, a.images, a.states)
    assert len(pool.get_batch(4).slots) == 4
