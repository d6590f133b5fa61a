"""Drop-in patching of a live reference model (Quanato607/XLSTM-HVED) onto the sm_100a kernels.

``patch_model(model)`` moves every reference ``ViLBlock`` and ``ProductOfExperts(2)`` instance found in ``model`` onto
a patched SUBCLASS of its own class (``instance.__class__`` is swapped; the subclass only overrides ``forward``) and
rebinds the module-level functions of the path -- ``reparametrize`` / ``clip`` (RA_HVED.py:741-753, looked up at
RA_HVED.py:597, 580), ``compute_KLD`` (loss.py:85) and the ``ZeroLayerF`` autograd function
(buildingblocks.py:308-323) -- in EVERY loaded module that holds a reference to the original object (``train.py:23``
and ``Pretrain.py:23`` do ``from loss import compute_KLD``, so rebinding ``loss.compute_KLD`` alone would not reach the
training loop).

Why a subclass and not an instance attribute:
  * ``nn.DataParallel`` (train.py:148-151) replicates modules with ``replica.__dict__ = module.__dict__.copy()``: a bound
    method stored on the instance would keep pointing at the device-0 module; a class-level ``forward`` resolves ``self``
    to the replica and therefore to the replica's parameters and device.
  * ``torch.save({'model': model})`` (train.py:370-397) pickles whole modules: the patched subclasses reduce themselves
    as their ORIGINAL class, so a checkpoint written from a patched model is byte-compatible with the reference (it
    loads without this package and comes back unpatched).
Parameters and state_dict keys are untouched; ``unpatch_model`` restores classes and globals exactly.
"""
from __future__ import annotations

import copyreg
import sys

import torch.nn as nn

from . import modules

_PATCHED = {}          # (original class, kind) -> patched subclass
_GLOBALS = []          # (module, attribute name, original object) of every rebound global


def _reduce_as_base(self, protocol):
    """Pickle a patched module as an instance of the reference's own class (stdlib reconstructor + the instance state:
    nothing of this package is needed to load the checkpoint)."""
    state = object.__reduce_ex__(self, 2)[2]
    if isinstance(state, dict) and "fused_slope" in state:          # patch-time annotation of the conv-path norms
        state = {k: v for k, v in state.items() if k != "fused_slope"}
    return (copyreg._reconstructor, (type(self)._xhved_base, object, None), state)


def _vil_forward(self, x):
    return modules.vil_block_forward(self, x)


def _poe_forward(self, mu_list, logvar_list, mod_list, eps=1e-8):
    return modules.product_of_experts(mu_list, logvar_list, mod_list, eps)


def _poe2_forward(self, mu, logvar, drop, eps=1e-8):
    return modules.product_of_experts_drop(mu, logvar, drop, eps)


def _inorm_forward(self, input):
    return modules.instance_norm_forward(self, input)


def _bnorm_forward(self, input):
    return modules.batch_norm_forward(self, input)


def _identity_forward(self, input):
    return input


def _dwconv3_forward(self, input):
    return modules.depthwise_conv3_forward(self, input)


def _pwconv_forward(self, input):
    return modules.pointwise_conv_forward(self, input)


def _conv3_forward(self, input):
    return modules.dense_conv3_forward(self, input)


def _atten2_forward(self, seg_x, enc_x, recon_x=None):
    return modules.atten_module2_forward(self, seg_x, enc_x, recon_x)


_FORWARDS = {"ViLBlock": _vil_forward, "ProductOfExperts": _poe_forward, "ProductOfExperts2": _poe2_forward,
             "InstanceNorm3d": _inorm_forward, "BatchNorm3d": _bnorm_forward, "FusedAwayLeakyReLU": _identity_forward,
             "AttenModule2": _atten2_forward, "DepthwiseConv3d": _dwconv3_forward,
             "PointwiseConv3d": _pwconv_forward, "DenseConv3d": _conv3_forward}


def _patched_class(cls, kind):
    key = (cls, kind)
    if key not in _PATCHED:
        _PATCHED[key] = type(cls.__name__, (cls,), {"forward": _FORWARDS[kind], "_xhved_base": cls, "__reduce_ex__": _reduce_as_base,
                                                    "__module__": cls.__module__, "__qualname__": cls.__qualname__})
    return _PATCHED[key]


def _kind_of(m):
    if hasattr(type(m), "_xhved_base"):
        return None                                   # already patched
    name = type(m).__name__
    if name == "ViLBlock" and hasattr(m, "layer") and hasattr(m.layer, "mlstm_cell"):
        return "ViLBlock"
    if name in ("ProductOfExperts", "ProductOfExperts2") and type(m).__module__ != modules.__name__:
        return name
    if type(m).__module__ != modules.__name__:
        # the conv path's normalisation layers: exact torch classes only, the mirrors already run the kernels
        if type(m) is nn.InstanceNorm3d and not m.track_running_stats:
            return "InstanceNorm3d"
        if type(m) is nn.BatchNorm3d:
            return "BatchNorm3d"
        if type(m) is nn.Conv3d and modules.dwconv3_supported(m):
            return "DepthwiseConv3d"
        if type(m) is nn.Conv3d and modules.pwconv_supported(m):
            return "PointwiseConv3d"
        if type(m) is nn.Conv3d and modules.conv3_supported(m):
            return "DenseConv3d"
        if (name == "AttenModule2" and all(hasattr(m, a) for a in ("compress", "enc_spatial", "enc_spatial2", "seg_spatial", "seg_spatial2"))
                and modules.gate_convs_supported(m.enc_spatial, m.enc_spatial2) and modules.gate_convs_supported(m.seg_spatial, m.seg_spatial2)):
            return "AttenModule2"
    return None


def _fuse_activations(model):
    """InstanceNorm3d / BatchNorm3d immediately followed by a LeakyReLU -- consecutive children of an nn.Sequential (SingleConv
    order 'ilc' / 'cil', buildingblocks.py:400-462; discriminator_block, buildingblocks.py:345-357) or the norm / relu pair of
    BasicConv (buildingblocks.py:11-31): the slope moves into the norm layer's kernel and the LeakyReLU hands its input through.
    Returns the number of fused pairs."""
    n = 0
    # a module instance registered in more than one place (a shared activation, a norm reused by two branches) is left alone:
    # folding would change what its other users compute
    uses = {}
    for parent in model.modules():
        for kid in parent._modules.values():
            if kid is not None:
                uses[id(kid)] = uses.get(id(kid), 0) + 1
    for parent in model.modules():
        if not (isinstance(parent, nn.Sequential) or type(parent).__name__ == "BasicConv"):
            continue
        kids = list(parent._modules.values())
        for norm, act in zip(kids, kids[1:]):
            if norm is None or act is None or getattr(type(norm), "_xhved_base", None) not in (nn.InstanceNorm3d, nn.BatchNorm3d):
                continue
            if type(act) is not nn.LeakyReLU or norm.__dict__.get("fused_slope") is not None:
                continue
            if uses.get(id(norm), 0) != 1 or uses.get(id(act), 0) != 1:
                continue
            norm.fused_slope = float(act.negative_slope)
            act.__class__ = _patched_class(nn.LeakyReLU, "FusedAwayLeakyReLU")
            n += 1
    return n


def _rebind_everywhere(original, replacement, attr):
    """Replace ``attr`` in every loaded module where it IS ``original`` (from-imports included).  Returns the count."""
    n = 0
    for mod in list(sys.modules.values()):
        d = getattr(mod, "__dict__", None)
        if d is None or d.get(attr) is not original:
            continue
        _GLOBALS.append((mod, attr, original))
        setattr(mod, attr, replacement)
        n += 1
    return n


def patch_model(model, patch_globals: bool = True, conv_path: bool = True):
    """Returns a dict with the number of patched objects per kind; ``globals`` counts (module, name) bindings rebound and
    ``rebound`` lists them as "module.name".  ``conv_path=False`` leaves the normalisation layers of the convolution path
    (InstanceNorm3d / BatchNorm3d and the LeakyReLU fused into them) on PyTorch."""
    counts = {"ViLBlock": 0, "ProductOfExperts": 0, "ProductOfExperts2": 0, "InstanceNorm3d": 0, "BatchNorm3d": 0,
              "AttenModule2": 0, "DepthwiseConv3d": 0, "PointwiseConv3d": 0, "DenseConv3d": 0, "fused_LeakyReLU": 0, "globals": 0, "rebound": []}
    for m in model.modules():
        kind = _kind_of(m)
        if kind in ("InstanceNorm3d", "BatchNorm3d", "AttenModule2", "DepthwiseConv3d", "PointwiseConv3d", "DenseConv3d") and not conv_path:
            continue
        if kind is not None:
            m.__class__ = _patched_class(type(m), kind)
            counts[kind] += 1
    if conv_path:
        counts["fused_LeakyReLU"] = _fuse_activations(model)
    if patch_globals:
        before = len(_GLOBALS)
        ra, ls, bb = sys.modules.get("RA_HVED"), sys.modules.get("loss"), sys.modules.get("buildingblocks")
        vl = next((m for n, m in sys.modules.items() if n.endswith("nets.vision_lstm") and hasattr(m, "parallel_stabilized_simple")), None)
        for owner, attr, repl in ((ra, "reparametrize", modules.reparametrize), (ra, "clip", modules.clip),
                                  (ls, "compute_KLD", modules.compute_KLD), (bb, "ZeroLayerF", modules.ZeroLayerF),
                                  # DiceLoss.forward calls this through the loss module's globals (loss.py:198-199)
                                  (ls, "compute_per_channel_dice", modules.compute_per_channel_dice),
                                  # the cell itself (vision_lstm.py:48-130, called at 327): reached by blocks wider than the
                                  # fused kernels cover and by any stand-alone MatrixLSTMCell of the reference
                                  (vl, "parallel_stabilized_simple", modules.parallel_stabilized_simple)):
            orig = getattr(owner, attr, None) if owner is not None else None
            if orig is None or orig is repl:
                continue
            _rebind_everywhere(orig, repl, attr)
        counts["globals"] = len(_GLOBALS) - before
        counts["rebound"] = [f"{mod.__name__}.{attr}" for mod, attr, _ in _GLOBALS[before:]]
    return counts


def unpatch_model(model):
    for m in model.modules():
        base = getattr(type(m), "_xhved_base", None)
        if base is not None:
            m.__class__ = base
            m.__dict__.pop("fused_slope", None)
    while _GLOBALS:
        mod, attr, orig = _GLOBALS.pop()
        setattr(mod, attr, orig)
