"""Import the read-only reference checkout unmodified (build container only).

Recipe from SURVEY.md Appendix A: four non-arithmetic imports are stubbed (easydict, skimage,
matplotlib, torchvision.transforms.functional_tensor); none of them is reached by any filter /
agent arithmetic.  ``/root/reference`` does not exist on the GPU box, so everything that uses this
module is either the golden generator (``tests/golden/make_golden.py``) or a test that skips itself
when ``available()`` is False.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("AISP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "isp", "filters.py"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


_loaded = None


def load():
    """Returns a namespace with the reference modules: filters, sharpen, denoise, agent, value, cfg."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF_ROOT}")
    import torch

    if "easydict" not in sys.modules:
        _stub("easydict", EasyDict=dict)
    if "skimage" not in sys.modules:
        sk = _stub("skimage")
        sk.io = _stub("skimage.io")
    if "matplotlib" not in sys.modules:
        mp = _stub("matplotlib")
        mp.pyplot = _stub("matplotlib.pyplot")
    _stub("torchvision.transforms.functional_tensor", torch_pad=torch.nn.functional.pad)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import isp.filters as rf
    import isp.sharpen as rs
    import isp.denoise as rd
    import agent as ra
    import value as rv
    from config import cfg

    _loaded = types.SimpleNamespace(filters=rf, sharpen=rs, denoise=rd, agent=ra, value=rv, cfg=cfg)
    return _loaded
