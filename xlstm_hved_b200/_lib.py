"""ctypes binding of libxhved.so (the C ABI declared in include/xhved.h).

There is NO CPU fallback: if the library is missing or no CUDA device is
present, every op raises.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_float, c_int, c_int64, c_uint32, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxhved.so")

_ERR = {-1: "XHVED_ERR_BAD_SHAPE", -2: "XHVED_ERR_UNSUPPORTED_DH", -3: "XHVED_ERR_BAD_ARG", -4: "XHVED_ERR_UNSUPPORTED_DIM"}

_PARAM_FIELDS = ["norm_weight", "proj_up_weight", "conv_weight", "conv_bias", "q_weight", "k_weight", "v_weight",
                 "igate_weight", "igate_bias", "fgate_weight", "fgate_bias", "outnorm_weight", "learnable_skip",
                 "proj_down_weight"]


class VilParams(Structure):
    _fields_ = [(n, c_void_p) for n in _PARAM_FIELDS]


class VilGrads(Structure):
    _fields_ = [(n, c_void_p) for n in _PARAM_FIELDS]


class VilShape(Structure):
    _fields_ = [("B", c_int), ("S", c_int), ("C", c_int), ("NH", c_int), ("QB", c_int), ("reverse", c_int),
                ("x_stride_b", c_int64), ("x_stride_n", c_int64), ("x_stride_c", c_int64),
                ("y_stride_b", c_int64), ("y_stride_n", c_int64), ("y_stride_c", c_int64),
                ("grad_replicas", c_int), ("grad_replica_stride", c_int64)]


class PoeLevel(Structure):            # xhved_poe_level
    _fields_ = [("mu", c_void_p), ("logvar", c_void_p), ("n", c_int64), ("expert_stride", c_int64), ("drop", c_void_p),
                ("per_sample", c_int64), ("out_mu", c_void_p), ("out_logvar", c_void_p), ("noise", c_void_p), ("out_z", c_void_p),
                ("kld_out", c_void_p)]


class PoeLevelGrad(Structure):        # xhved_poe_level_grad
    _fields_ = [("mu", c_void_p), ("logvar", c_void_p), ("n", c_int64), ("expert_stride", c_int64), ("drop", c_void_p),
                ("per_sample", c_int64), ("g_mu", c_void_p), ("g_logvar", c_void_p), ("noise", c_void_p), ("g_z", c_void_p),
                ("kld_scale", POINTER(c_float)), ("d_mu", c_void_p), ("d_logvar", c_void_p)]


class PoeOpts(Structure):             # xhved_poe_opts
    _fields_ = [("eps", c_float), ("flags", c_int), ("clip_lo", c_float), ("clip_hi", c_float)]


class NormShape(Structure):           # xhved_norm_shape
    _fields_ = [("N", c_int), ("C", c_int), ("spatial", c_int64), ("mode", c_int), ("dtype", c_int), ("eps", c_float), ("slope", c_float)]


class MlstmWorkspace(Structure):      # xhved_mlstm_workspace
    _fields_ = [("nc", c_int), ("dhp", c_int), ("tile_bytes", c_int64), ("row_bytes", c_int64), ("dstate_bytes", c_int64),
                ("chunk_bytes", c_int64), ("states_bytes", c_int64)]


class VilWorkspaceSizes(Structure):   # xhved_vil_workspace
    _fields_ = [("cell", MlstmWorkspace), ("token_tile_bytes", c_int64), ("grad_replica_stride", c_int64)]


# every symbol include/xhved.h declares -> argtypes (None = not yet bound with a signature)
SYMBOLS = {
    "xhved_version": [],
    "xhved_poe_fwd": [c_void_p, c_void_p, c_int64, c_int64, POINTER(c_uint32), c_int, c_void_p, c_int64, c_float, c_void_p,
                      c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "xhved_poe_bwd": [c_void_p, c_void_p, c_int64, c_int64, POINTER(c_uint32), c_int, c_void_p, c_int64, c_float, c_void_p,
                      c_void_p, c_void_p, c_void_p, POINTER(c_float), c_void_p, c_void_p, c_int, c_void_p],
    "xhved_poe_fwd_levels": [POINTER(PoeLevel), c_int, POINTER(c_uint32), c_int, c_float, c_int, c_void_p],
    "xhved_poe_bwd_levels": [POINTER(PoeLevelGrad), c_int, POINTER(c_uint32), c_int, c_float, c_int, c_void_p],
    "xhved_poe_fwd_levels_opts": [POINTER(PoeLevel), c_int, POINTER(c_uint32), c_int, POINTER(PoeOpts), c_void_p],
    "xhved_poe_bwd_levels_opts": [POINTER(PoeLevelGrad), c_int, POINTER(c_uint32), c_int, POINTER(PoeOpts), c_void_p],
    "xhved_clip_fwd": [c_void_p, c_int64, c_float, c_float, c_void_p, c_void_p],
    "xhved_clip_bwd": [c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p, c_void_p],
    "xhved_zero_rows": [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p],
    "xhved_dice_sums": [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p],
    "xhved_dice_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_float, c_void_p, c_void_p],
    "xhved_norm_act_workspace": [c_int, c_int, c_int64, c_int],
    "xhved_norm_act_fwd": [c_void_p, c_void_p, c_void_p, POINTER(NormShape), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "xhved_norm_act_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(NormShape), c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p],
    "xhved_gate7_workspace": [c_int] * 5,
    "xhved_gate7_fwd": [c_void_p, c_void_p, c_void_p] + [c_int] * 5 + [c_void_p, c_void_p],
    "xhved_gate7_bwd": [c_void_p] * 4 + [c_int] * 5 + [c_void_p] * 5,
    "xhved_dwconv3_workspace": [c_int] * 5,
    "xhved_dwconv3_fwd": [c_void_p, c_void_p, c_void_p] + [c_int] * 5 + [c_void_p, c_void_p],
    "xhved_dwconv3_bwd": [c_void_p] * 3 + [c_int] * 5 + [c_void_p] * 5,
    "xhved_pwconv_workspace": [c_int, c_int, c_int, c_int64],
    "xhved_pwconv_fwd": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int, c_void_p, c_void_p],
    "xhved_pwconv_bwd": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int] + [c_void_p] * 5,
    "xhved_conv3_workspace": [c_int] * 6,
    "xhved_conv3_fwd": [c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p, c_void_p],
    "xhved_conv3_bwd": [c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p] * 5,
    "xhved_reparam_fwd": [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p],
    "xhved_reparam_bwd": [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p],
    "xhved_mlstm_fwd": [c_void_p] * 5 + [c_int] * 4 + [c_float] + [c_void_p] * 9,
    "xhved_mlstm_bwd": [c_void_p] * 11 + [c_int] * 4 + [c_float] + [c_void_p] * 12,
    "xhved_mlstm_pack": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "xhved_mlstm_pack_gates": [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "xhved_mlstm_unpack": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "xhved_mlstm_workspace_query": [c_int, c_int, c_int, POINTER(MlstmWorkspace)],
    "xhved_vil_workspace_query": [c_int, c_int, c_int, POINTER(VilWorkspaceSizes)],
    "xhved_profile_enable": [c_int],
    "xhved_profile_kernel_count": [],
    "xhved_profile_kernel_name": [c_int],
    "xhved_profile_read": [POINTER(c_float), POINTER(c_int), c_int],
    "xhved_reduce_replicas": [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p],
    "xhved_umma_issue_bench": [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "xhved_umma_selftest": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "xhved_vil_wide_pre_fwd": [c_void_p] * 10 + [c_int] * 4 + [c_void_p] * 7,
    "xhved_vil_wide_post_fwd": [c_void_p] * 5 + [c_int] * 4 + [c_void_p] * 2,
    "xhved_vil_wide_post_bwd": [c_void_p] * 6 + [c_int] * 4 + [c_void_p] * 6,
    "xhved_vil_wide_pre_bwd": [c_void_p] * 8 + [c_int] * 4 + [c_void_p] * 14,
    "xhved_split_hilo_cat": [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p],
    "xhved_vil_block_workspace": [c_int, c_int, c_int, c_int, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)],
    "xhved_vil_block_fwd": [c_void_p, POINTER(VilParams), POINTER(VilShape), c_float, c_void_p, c_void_p, c_void_p],
    "xhved_vil_block_bwd": [c_void_p, c_void_p, POINTER(VilParams), POINTER(VilShape), c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p],
    "xhved_vil_pre_fwd": [c_void_p, POINTER(VilParams), POINTER(VilShape)] + [c_void_p] * 9,
    "xhved_vil_post_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(VilParams), POINTER(VilShape), c_void_p, c_void_p],
    "xhved_vil_post_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(VilParams), POINTER(VilShape), c_void_p, c_void_p,
                           c_void_p, POINTER(VilGrads), c_void_p],
    "xhved_vil_pre_bwd": [c_void_p] * 13 + [POINTER(VilParams), POINTER(VilShape), c_void_p, POINTER(VilGrads), c_void_p, c_void_p, c_void_p],
}

_lib = None


def load_library() -> ctypes.CDLL:
    """dlopen libxhved.so and bind every declared symbol.  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m xlstm_hved_b200.build` "
                "(this package has no CPU or PyTorch fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        missing = []
        for name, argtypes in SYMBOLS.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                missing.append(name)
                continue
            fn.argtypes = argtypes
            fn.restype = (ctypes.c_char_p if name == "xhved_profile_kernel_name" else
                          c_int64 if name in ("xhved_norm_act_workspace", "xhved_gate7_workspace", "xhved_dwconv3_workspace", "xhved_pwconv_workspace", "xhved_conv3_workspace") else c_int)
        if missing:
            raise RuntimeError(f"{LIB_PATH} lacks symbols declared in include/xhved.h: {missing}; rebuild it")
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError(f"{what}: {_ERR.get(rc, rc)}")
    raise RuntimeError(f"{what}: CUDA error {rc} ({torch.cuda.get_device_name() if torch.cuda.is_available() else 'no device'})")


def ptr(t):
    """Device pointer of a tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("xlstm_hved_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()          # a plain int: ctypes converts it for the c_void_p parameter (0 would be NULL: never a live tensor)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream():
    """cudaStream_t of the calling thread's current stream on the current device, as an integer (None = the legacy default)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device()) or None
    return torch.cuda.current_stream().cuda_stream or None


def on_device(fn):
    """Run an autograd.Function forward / backward with the CUDA device of its first CUDA tensor argument current.
    The kernels launch on the calling thread's current device and stream; under nn.DataParallel (train.py:148-151) every
    replica runs in its own thread, and autograd's backward threads do not inherit a ``torch.cuda.device`` context."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():      # the common case: no context switch (~10 us per call)
                    return fn(*args, **kwargs)
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapped
