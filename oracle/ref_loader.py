"""Import the real reference (Quanato607/XLSTM-HVED) for validation  --  TEST INFRASTRUCTURE.

Looks for the reference tree at $XHVED_REFERENCE, /root/reference, then
<repo>/baseline/_ref.  The reference is pure Python; a few third-party modules
it imports at module scope (none of which hold arithmetic used on this path)
are absent from the image and are registered as empty stubs (SURVEY.md
Appendix B).  The GPU box has no reference tree: callers must handle
``find_reference() is None`` (tests skip; golden fixtures are used instead).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference():
    for cand in (os.environ.get("XHVED_REFERENCE"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "RA_HVED.py")):
            return cand
    return None


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_stub(parent), child, mod)
    return mod


def _install_stubs():
    noop = lambda *a, **k: None
    empty = type("Stub", (), {})
    _stub("dynamic_network_architectures.building_blocks.helper",
          get_matching_convtransp=noop, convert_conv_op_to_dim=noop, get_matching_instancenorm=noop,
          convert_dim_to_conv_op=noop, maybe_convert_scalar_to_list=noop, get_matching_pool_op=noop)
    _stub("dynamic_network_architectures.initialization.weight_init", init_last_bn_before_add_to_0=noop)
    _stub("dynamic_network_architectures.building_blocks.residual", BasicBlockD=empty)
    _stub("nnunetv2.utilities.plans_handling.plans_handler", ConfigurationManager=empty, PlansManager=empty)
    _stub("nnunetv2.utilities.network_initialization", InitWeights_He=empty)
    for name in ("h5py", "skimage", "skimage.segmentation"):
        try:
            __import__(name)
        except Exception:
            _stub(name, find_boundaries=noop)
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        _stub("matplotlib.pyplot", ioff=noop, switch_backend=noop)


_loaded = {}


def load_vision_lstm():
    """The ViL module alone (needs no stubs): UxLSTM/nnunetv2/nets/vision_lstm.py."""
    ref = find_reference()
    if ref is None:
        return None
    if "vl" not in _loaded:
        sys.dont_write_bytecode = True
        if ref not in sys.path:
            sys.path.insert(0, ref)
        import importlib
        _loaded["vl"] = importlib.import_module("UxLSTM.nnunetv2.nets.vision_lstm")
    return _loaded["vl"]


def load_reference():
    """Returns a namespace with RA_HVED, buildingblocks, loss, utils, vision_lstm,
    UxLSTMEnc_3d of the real reference, or None if no reference tree exists."""
    ref = find_reference()
    if ref is None:
        return None
    if "all" in _loaded:
        return _loaded["all"]
    import torch
    sys.dont_write_bytecode = True
    if ref not in sys.path:
        sys.path.insert(0, ref)
    _install_stubs()
    if not torch.cuda.is_available():
        # RA_HVED.py:520 calls .cuda() on an (unused) mask tensor unconditionally
        torch.Tensor.cuda = lambda self, *a, **k: self
    import importlib
    with contextlib.redirect_stdout(io.StringIO()):
        ns = types.SimpleNamespace(
            RA_HVED=importlib.import_module("RA_HVED"),
            buildingblocks=importlib.import_module("buildingblocks"),
            loss=importlib.import_module("loss"),
            utils=importlib.import_module("utils"),
            vision_lstm=importlib.import_module("UxLSTM.nnunetv2.nets.vision_lstm"),
            UxLSTMEnc_3d=importlib.import_module("UxLSTM.nnunetv2.nets.UxLSTMEnc_3d"),
        )
    _loaded["all"] = ns
    return ns


def build_model(f_maps: int = 4, seed: int = 1):
    """XLSTM_HVED exactly as train.py:142-145 builds it (seeded, init_weights applied)."""
    ns = load_reference()
    if ns is None:
        return None
    with contextlib.redirect_stdout(io.StringIO()):
        ns.utils.seed_everything(seed)
        model = ns.RA_HVED.XLSTM_HVED(1, 3, multi_stream=4, fusion_level=4, shared_recon=True, recon_skip=True,
                                     MVAE_reduction=True, final_sigmoid=True, f_maps=f_maps, layer_order="ilc")
        model.apply(ns.utils.init_weights)
    return model
