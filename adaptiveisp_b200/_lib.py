"""ctypes binding of ``libaisp_b200.so`` (the C ABI declared in ``include/aisp_b200.h``).

There is no fallback of any kind: if the shared library is missing the first call raises with the
build command, and tensors that are not CUDA / fp32 / contiguous are rejected, never converted.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_float, c_int, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libaisp_b200.so")

PSTRIDE = 24
MAX_STEPS = 8
MAX_CHAIN_BWD = 6
MAX_BANK_FILTERS = 16
SEQ_CLIP, SEQ_STRICT = 1, 2   # bits of `clip_each` in the sequence entry points
ABI_VERSION = 3

# name -> (restype, argtypes); mirrors include/aisp_b200.h one to one
_P = c_void_p
SIGNATURES = {
    "aisp_version": (c_int, []),
    "aisp_status_string": (c_char_p, [c_int]),
    "aisp_op_num_params": (c_int, [c_int]),
    "aisp_bwd_scratch_bytes": (c_size_t, [c_int, c_int, c_int]),
    "aisp_pointwise_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "aisp_pointwise_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "aisp_pointwise_chain_bwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, c_size_t,
                                         _P]),
    "aisp_sharpen_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "aisp_sharpen_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "aisp_nlm_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "aisp_nlm_module_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, c_int, _P]),
    "aisp_shot_read_noise": (c_int, [_P, _P, _P, _P, _P, _P, c_int, ctypes.c_longlong, ctypes.c_ulonglong,
                                     ctypes.c_ulonglong, _P]),
    "aisp_nlm_param_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "aisp_nlm_bwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "aisp_nlm_bwd_img": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "aisp_block_mean": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "aisp_value_stats": (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    "aisp_select_apply_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "aisp_select_apply_bwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, c_size_t,
                                      _P]),
    "aisp_sequence_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, c_int, _P, c_int,
                                  c_int, _P, _P, _P]),
    "aisp_select_apply_bwd_pooled": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P,
                                             _P, _P, c_size_t, _P]),
    "aisp_select": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, c_int, c_int, c_int, c_float, c_float, _P, _P, _P, _P, _P,
                            _P, _P]),
    "aisp_select_bwd": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "aisp_regress_fwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "aisp_regress_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "aisp_bank_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "aisp_bank_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, c_size_t, _P]),
}

_lib = None


class AispError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the library in place for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j4"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise AispError("building libaisp_b200.so failed (see output above)")
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise AispError(
                f"{LIB_PATH} not found: the CUDA library is required (no CPU fallback). "
                f"Build it with `make -C {CSRC}` or `python -c 'import __graft_entry__ as g; g.build()'`.")
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)  # AttributeError here == header and library disagree
            fn.restype = res
            fn.argtypes = args
        v = h.aisp_version()
        if v != ABI_VERSION:
            raise AispError(f"libaisp_b200.so ABI version {v}, binding expects {ABI_VERSION}")
        _lib = h
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().aisp_status_string(status).decode()
        raise AispError(f"{what} failed: {msg} (status {status})")


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def require_image(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise AispError(f"{name} must be a CUDA tensor: the ISP kernels have no CPU path")
    if t.dtype != torch.float32:
        raise AispError(f"{name} must be float32, got {t.dtype}")
    if t.dim() != 4 or t.shape[1] != 3:
        raise AispError(f"{name} must be [B,3,H,W], got {tuple(t.shape)}")
    if not t.is_contiguous():
        raise AispError(f"{name} must be contiguous NCHW")


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def scratch(B: int, H: int, W: int, device: torch.device) -> torch.Tensor:
    """Partial-sum buffer for one backward call (``aisp_bwd_scratch_bytes``).

    Allocated per call from the caching allocator: that costs no cudaMalloc in steady state and is
    safe under CUDA-graph capture (the block comes from the graph's private pool and lives as long
    as the graph does) -- a cached grow-only buffer would be dropped on growth while a captured
    graph still writes to it on replay."""
    need = lib().aisp_bwd_scratch_bytes(B, H, W)
    return torch.empty(max(need, 1 << 12), dtype=torch.uint8, device=device)
