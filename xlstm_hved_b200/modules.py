"""Host-side mirror of the reference's module API for the hot path (same class names, constructor arguments,
parameter names / state_dict keys and forward signatures), backed by the sm_100a kernels.

Reference classes mirrored (paths relative to the reference root):
  UxLSTM/nnunetv2/nets/vision_lstm.py : SequenceTraversal (15-17), LinearHeadwiseExpand (133-175), CausalConv1d
      (178-221), LayerNorm (224-268), MultiHeadLayerNorm (271-287), MatrixLSTMCell (290-348), ViLLayer (351-477),
      ViLBlock (480-506), parallel_stabilized_simple (48-130)
  UxLSTM/nnunetv2/nets/UxLSTMEnc_3d.py : ViLLayer (42-87)  -> here ``ViLLayer3D``
  buildingblocks.py : ProductOfExperts (846-866), ProductOfExperts2 (868-886)
  RA_HVED.py        : reparametrize (741-747), clip (749-753), SUBSETS_MODALITIES (733-738)
  loss.py           : KL_divergence (29-40), compute_KLD (85-115)

The containers below own the parameters exactly as the reference does, so ``load_state_dict(strict=True)``
round-trips checkpoints of the reference modules.  The model path runs only through the fused entry points
(ViLBlock / ViLLayer3D / parallel_stabilized_simple / ProductOfExperts(2) / reparametrize / compute_KLD).

The sub-modules also keep the reference's stand-alone forwards (SURVEY 8b entry points 3 and 4: inner ``ViLLayer`` and
``MatrixLSTMCell``): called on their own they run the S x S-shaped work -- the mLSTM cell -- through the same
tcgen05 kernels (``xhved_mlstm_fwd/bwd``) and only the per-token glue around it (a (3E -> 4) gate Linear, the 4-tap conv,
the 4x4 block projections, the norms) as device-side torch ops.  They need CUDA tensors and the built library like
everything else here: there is no CPU path, and a fused ``ViLBlock`` never calls them.
"""
from __future__ import annotations

import math
from enum import Enum

import torch
from torch import nn

import torch.nn.functional as F
from torch.autograd.function import once_differentiable

from . import _lib, ops
from .ops import SUBSETS_MODALITIES, parallel_stabilized_simple  # noqa: F401  (re-exported drop-in)


class SequenceTraversal(Enum):
    ROWWISE_FROM_TOP_LEFT = "rowwise_from_top_left"
    ROWWISE_FROM_BOT_RIGHT = "rowwise_from_bot_right"


def _is_reverse(direction) -> bool:
    value = getattr(direction, "value", direction)
    if value == SequenceTraversal.ROWWISE_FROM_TOP_LEFT.value:
        return False
    if value == SequenceTraversal.ROWWISE_FROM_BOT_RIGHT.value:
        return True
    raise NotImplementedError(direction)       # vision_lstm.py:423-424


def _require_device(*tensors):
    """Stand-alone sub-module forwards are device-only, like the fused entry points."""
    _lib.load_library()                                  # raises if libxhved.so has not been built
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("xlstm_hved_b200 has no CPU path")


class _Holder(nn.Module):
    pass


class LinearHeadwiseExpand(_Holder):
    def __init__(self, dim, num_heads, bias=False):
        super().__init__()
        assert dim % num_heads == 0 and not bias
        self.dim, self.num_heads = dim, num_heads
        d = dim // num_heads
        self.weight = nn.Parameter(torch.empty(num_heads, d, d))
        self.bias = None
        nn.init.normal_(self.weight.data, mean=0.0, std=math.sqrt(2 / 5 / d))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Block-diagonal linear (vision_lstm.py:159-165): every group of d channels is multiplied by its own d x d block."""
        _require_device(x)
        lead = x.shape[:-1]
        d = self.dim // self.num_heads
        xb = x.reshape(*lead, self.num_heads, 1, d)
        # broadcast multiply + reduce instead of an einsum: the einsum lowers to a batched GEMM over 4 x 4 blocks, which costs
        # 0.7 ms per call at (16, 4096, 256) where this elementwise form takes 0.1 ms
        return (xb * self.weight).sum(-1).reshape(*lead, self.dim)


class CausalConv1d(_Holder):
    def __init__(self, dim, kernel_size=4, bias=True):
        super().__init__()
        assert kernel_size == 4 and bias, "the fused kernel implements the reference's k=4, bias=True conv"
        self.dim, self.kernel_size, self.bias, self.pad = dim, kernel_size, bias, kernel_size - 1
        self.conv = nn.Conv1d(dim, dim, kernel_size=kernel_size, padding=self.pad, groups=dim, bias=bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Depthwise conv over the 3 previous tokens and the current one (vision_lstm.py:213-221); x: (B, S, dim)."""
        _require_device(x)
        # four shifted multiply-adds on the (B, S, dim) layout: no transposes, no depthwise cuDNN convolution
        xp = F.pad(x, (0, 0, self.pad, 0))                        # left padding over tokens only == causal
        S = x.shape[1]
        w = self.conv.weight[:, 0, :]                             # (dim, 4): tap j multiplies token t - 3 + j
        y = self.conv.bias + xp[:, 0:S] * w[:, 0]
        for j in range(1, self.kernel_size):
            y = y + xp[:, j:j + S] * w[:, j]
        return y


class LayerNorm(_Holder):
    def __init__(self, ndim=-1, weight=True, bias=False, eps=1e-5, residual_weight=True):
        super().__init__()
        assert weight and not bias and residual_weight and eps == 1e-5
        self.weight = nn.Parameter(torch.zeros(ndim))
        self.bias = None
        self.eps, self.residual_weight, self.ndim = eps, residual_weight, ndim

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """vision_lstm.py:258-268: layer norm over the last dim with weight 1 + w and no bias."""
        _require_device(x)
        return F.layer_norm(x, (self.ndim,), weight=1.0 + self.weight, bias=None, eps=self.eps)


class MultiHeadLayerNorm(LayerNorm):
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """vision_lstm.py:271-287: x (B, NH, S, DH) -> (B, S, NH*DH), every (token, head) normalised over DH."""
        _require_device(x)
        B, NH, S, DH = x.shape
        xt = x.transpose(1, 2)                                      # (B, S, NH, DH)
        xc = xt - xt.mean(dim=-1, keepdim=True)
        xhat = xc * torch.rsqrt(xc.pow(2).mean(dim=-1, keepdim=True) + self.eps)
        return (xhat * (1.0 + self.weight).view(NH, DH)).reshape(B, S, NH * DH)


class MatrixLSTMCell(_Holder):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.dim, self.num_heads = dim, num_heads
        self.igate = nn.Linear(3 * dim, num_heads)
        self.fgate = nn.Linear(3 * dim, num_heads)
        self.outnorm = MultiHeadLayerNorm(ndim=dim, weight=True, bias=False)
        self.causal_mask_cache = {}
        self.reset_parameters()

    def reset_parameters(self):                      # vision_lstm.py:341-348
        nn.init.zeros_(self.fgate.weight)
        with torch.no_grad():
            self.fgate.bias.copy_(torch.linspace(3.0, 6.0, self.fgate.bias.shape[0]))
        nn.init.zeros_(self.igate.weight)
        nn.init.normal_(self.igate.bias, mean=0.0, std=0.1)

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
        """vision_lstm.py:302-339: gates from [q|k|v], heads split off the channel dim, stabilised cell, per-head norm.
        The cell runs on the tcgen05 kernels (causal by construction: no S x S mask is built)."""
        _require_device(q, k, v)
        B, S, _ = q.shape
        nh = self.num_heads
        qkv = torch.cat([q, k, v], dim=-1)
        ig = F.linear(qkv, self.igate.weight, self.igate.bias).transpose(1, 2).unsqueeze(-1)     # (B, NH, S, 1)
        fg = F.linear(qkv, self.fgate.weight, self.fgate.bias).transpose(1, 2).unsqueeze(-1)
        heads = lambda t: t.reshape(B, S, nh, -1).transpose(1, 2)                                 # (B, NH, S, DH)
        h = ops.parallel_stabilized_simple(heads(q), heads(k), heads(v), ig, fg)
        return self.outnorm(h)


class ViLLayer(_Holder):
    """Inner layer (vision_lstm.py:351-477): owns the projections, conv, cell and skip parameters."""

    def __init__(self, dim, direction, expansion=2, qkv_block_size=4, proj_bias=False, conv_bias=True, kernel_size=4):
        super().__init__()
        if dim % qkv_block_size != 0:
            qkv_block_size = 2
        assert expansion == 2 and not proj_bias and conv_bias and kernel_size == 4
        self.dim, self.direction, self.expansion, self.qkv_block_size = dim, direction, expansion, qkv_block_size
        self.proj_bias, self.conv_bias, self.kernel_size = proj_bias, conv_bias, kernel_size
        inner = expansion * dim
        nh = inner // qkv_block_size
        self.proj_up = nn.Linear(dim, 2 * inner, bias=False)
        self.q_proj = LinearHeadwiseExpand(inner, nh)
        self.k_proj = LinearHeadwiseExpand(inner, nh)
        self.v_proj = LinearHeadwiseExpand(inner, nh)
        self.conv1d = CausalConv1d(inner, kernel_size=kernel_size, bias=conv_bias)
        self.mlstm_cell = MatrixLSTMCell(inner, num_heads=qkv_block_size)
        self.learnable_skip = nn.Parameter(torch.ones(inner))
        self.proj_down = nn.Linear(inner, dim, bias=False)
        self.reset_parameters()

    def reset_parameters(self):                      # vision_lstm.py:455-477
        nn.init.normal_(self.proj_up.weight, mean=0.0, std=math.sqrt(2 / (5 * self.dim)))
        nn.init.normal_(self.proj_down.weight, mean=0.0, std=2 / math.sqrt(self.dim))
        nn.init.ones_(self.learnable_skip)
        for proj in (self.q_proj, self.k_proj, self.v_proj):
            nn.init.normal_(proj.weight, mean=0.0, std=math.sqrt(2 / (5 * self.dim)))
        self.mlstm_cell.reset_parameters()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Stand-alone inner layer (vision_lstm.py:415-453), x: (B, S, dim).  Inside a ViLBlock this is never called -- the
        block runs the fused pre / cell / post kernels; here only the cell is a custom kernel."""
        _require_device(x)
        rev = _is_reverse(self.direction)
        if rev:
            x = x.flip(dims=[1])
        x_mlstm, z = F.linear(x, self.proj_up.weight).chunk(2, dim=-1)
        act = F.silu(self.conv1d(x_mlstm))
        h = self.mlstm_cell(self.q_proj(act), self.k_proj(act), self.v_proj(x_mlstm))
        y = F.linear((h + self.learnable_skip * act) * F.silu(z), self.proj_down.weight)
        return y.flip(dims=[1]) if rev else y


class DropPath(nn.Sequential):
    """vision_lstm_util.py:133-209 with the only configuration the hot path uses (drop_prob = 0)."""

    def __init__(self, *args, drop_prob: float = 0.0, scale_by_keep: bool = True, stochastic_drop_prob: bool = False):
        super().__init__(*args)
        assert 0.0 <= drop_prob < 1.0
        self._drop_prob = drop_prob
        self.scale_by_keep, self.stochastic_drop_prob = scale_by_keep, stochastic_drop_prob

    @property
    def drop_prob(self):
        return self._drop_prob


def vil_block_params(block):
    """The 14 parameter tensors of a (reference or mirror) ViLBlock in kernel order (ops.VIL_PARAM_KEYS)."""
    lay, cell = block.layer, block.layer.mlstm_cell
    return [block.norm.weight, lay.proj_up.weight, lay.conv1d.conv.weight, lay.conv1d.conv.bias, lay.q_proj.weight,
            lay.k_proj.weight, lay.v_proj.weight, cell.igate.weight, cell.igate.bias, cell.fgate.weight, cell.fgate.bias,
            cell.outnorm.weight, lay.learnable_skip, lay.proj_down.weight]


FUSED_DIMS = (16, 32, 64)        # model dims the fused K2 / K3 kernels are built for (f_maps 2 / 4 / 8: vil_pre.cu, vil_post.cu)


def vil_block_forward(block, x: torch.Tensor) -> torch.Tensor:
    """ViLBlock.forward (vision_lstm.py:499-502): x + layer(norm(x)) for a (B,S,C) token tensor or view.

    dims 16 / 32 / 64: the fused pre / cell / post kernels.  dims 128 / 256 (f_maps 16 / 32, head dim 64 / 128, SURVEY 8d
    config 2 (iii)): ``ops.vil_block_wide`` -- the three Linear layers as library GEMMs, the cell on the tcgen05 kernels and
    everything between them on the fused glue kernels of csrc/vil_wide.cu.  Any other width: the cell on the kernels, the
    per-token glue as device-side torch ops (the mirror modules' own forwards, or, for a patched reference block, the
    reference's own ``forward`` with ``parallel_stabilized_simple`` rebound to the kernels, patch.py)."""
    _require_device(x)
    dp = getattr(block.drop_path, "drop_prob", 0.0)
    if dp != 0.0 and block.training:
        raise NotImplementedError("stochastic depth (drop_path > 0 in training) is never used by XLSTM-HVED")
    if block.layer.qkv_block_size != 4 or getattr(block.norm, "bias", None) is not None:
        raise NotImplementedError("fused ViL block: qkv_block_size must be 4 and the norm bias-free (reference defaults)")
    if x.shape[-1] in FUSED_DIMS:
        return ops.vil_block(x, vil_block_params(block), reverse=_is_reverse(block.direction))
    if x.shape[-1] in ops.WIDE_DIMS:                    # the three Linear layers as library GEMMs, the rest fused (csrc/vil_wide.cu)
        return ops.vil_block_wide(x, vil_block_params(block), reverse=_is_reverse(block.direction))
    stock = getattr(type(block), "_xhved_base", None)
    if stock is not None:                               # a patched reference block: its own glue, our cell
        return stock.forward(block, x)
    return x + block.layer(block.norm(x))


class ViLBlock(nn.Module):
    def __init__(self, dim, direction, drop_path=0.0, norm_bias=False):
        super().__init__()
        assert not norm_bias
        self.dim, self.direction, self.norm_bias = dim, direction, norm_bias
        self.drop_path = DropPath(drop_prob=drop_path)
        self.norm = LayerNorm(ndim=dim, weight=True, bias=norm_bias)
        self.layer = ViLLayer(dim=dim, direction=direction)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return vil_block_forward(self, x)


def vil_wrapper_forward(wrapper, x: torch.Tensor) -> torch.Tensor:
    """Outer 3-D ViLLayer.forward (UxLSTMEnc_3d.py:54-63, 77-87): NCDHW feature -> token view -> block -> NCDHW.
    No transposed copies are made: the kernels read and write the strided token view in place."""
    if wrapper.channel_token:
        raise NotImplementedError("channel_token=True is never used by XLSTM-HVED (RA_HVED.py:314)")
    with torch.autocast(device_type="cuda", enabled=False):
        if x.dtype != torch.float32:
            x = x.float()                                  # the reference up-casts fp16 (UxLSTMEnc_3d.py:78-79)
        B, d_model = x.shape[:2]
        assert d_model == wrapper.dim
        img_dims = x.shape[2:]
        x = x.contiguous()
        x_flat = x.reshape(B, d_model, -1).transpose(-1, -2)
        x_vil = wrapper.vil(x_flat)
        return x_vil.transpose(-1, -2).reshape(B, d_model, *img_dims)


class ViLLayer3D(nn.Module):
    """Mirror of UxLSTMEnc_3d.ViLLayer (42-87).  ``norm`` is never used in forward but owns two parameters that
    must stay in the state_dict (UxLSTMEnc_3d.py:47)."""

    def __init__(self, dim, d_state=16, d_conv=4, expand=2, channel_token=False):
        super().__init__()
        self.dim = dim
        self.norm = nn.LayerNorm(dim)
        self.vil = ViLBlock(dim=dim, direction=SequenceTraversal.ROWWISE_FROM_TOP_LEFT)
        self.channel_token = channel_token

    def forward(self, x):
        return vil_wrapper_forward(self, x)


# --------------------------------------------------------------------------------------------- S-MVAE
class _PoEFunction(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, mu5, logvar5, subset, drop, eps):
        pm, pl, _, _ = ops.poe_fwd(mu5, logvar5, [subset], drop=drop, eps=eps)
        ctx.save_for_backward(mu5, logvar5)
        ctx.subset, ctx.drop, ctx.eps = subset, drop, eps
        return pm[0], pl[0]

    @staticmethod
    @_lib.on_device
    def backward(ctx, g_mu, g_lv):
        mu5, logvar5 = ctx.saved_tensors
        d_mu, d_lv = ops.poe_bwd(mu5, logvar5, [ctx.subset], g_mu=g_mu.contiguous()[None], g_logvar=g_lv.contiguous()[None],
                                 drop=ctx.drop, eps=ctx.eps)
        return d_mu, d_lv, None, None, None


def _stack5(t):
    if isinstance(t, (list, tuple)):
        t = torch.stack(list(t), 0)
    if t.shape[0] != 5:
        raise ValueError("expected (5, B, ...) experts with the prior at index 0 (RA_HVED.py:576-580)")
    return t.float().contiguous()


def product_of_experts(mu_list, logvar_list, mod_list, eps=1e-8):
    """ProductOfExperts.forward (buildingblocks.py:853-866)."""
    return _PoEFunction.apply(_stack5(mu_list), _stack5(logvar_list), tuple(int(m) for m in mod_list), None, eps)


def product_of_experts_drop(mu, logvar, drop, eps=1e-8):
    """ProductOfExperts2.forward (buildingblocks.py:875-886), including its in-place zeroing of ``mu``.

    The reference overwrites mu[m+1] with a copy whose dropped samples are zero (879-880).  The same side effect is
    applied here BEFORE the fused op runs (so the tensor autograd saves is the final one); the kernel removes the dropped
    experts through the mask itself and returns zero gradients for them, which is what ZeroLayerF's backward does."""
    if isinstance(mu, torch.Tensor):
        with torch.no_grad():
            for m in range(drop.shape[1]):
                mu[m + 1][drop[:, m].bool()] = 0
    mu5, lv5 = _stack5(mu), _stack5(logvar)
    return _PoEFunction.apply(mu5, lv5, (0, 1, 2, 3), drop.to(torch.uint8).contiguous(), eps)


class ProductOfExperts(nn.Module):
    def forward(self, mu_list, logvar_list, mod_list, eps=1e-8):
        return product_of_experts(mu_list, logvar_list, mod_list, eps)


class ProductOfExperts2(nn.Module):
    def forward(self, mu, logvar, drop, eps=1e-8):
        return product_of_experts_drop(mu, logvar, drop, eps)


class _ReparamFunction(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, mu, logvar, noise):
        ctx.save_for_backward(logvar, noise)
        return ops.reparam_fwd(mu, logvar, noise)

    @staticmethod
    @_lib.on_device
    def backward(ctx, g):
        logvar, noise = ctx.saved_tensors
        d_mu, d_lv = ops.reparam_bwd(logvar, noise, g.contiguous())
        return d_mu, d_lv, None


def reparametrize(mu, logvar, valid=False):
    """RA_HVED.py:741-747.  The noise comes from the same torch call, generator, shape and order as the
    reference's ``std.data.new(std.size()).normal_()`` so seeded runs draw identical samples."""
    if valid:
        return mu
    noise = torch.empty_like(mu, dtype=torch.float32).normal_()
    return _ReparamFunction.apply(mu.float().contiguous(), logvar.float().contiguous(), noise)


class _ClipFunction(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, x, lo, hi):
        ctx.save_for_backward(x)
        ctx.lo, ctx.hi = lo, hi
        return ops.clip_fwd(x, lo, hi)

    @staticmethod
    @_lib.on_device
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ops.clip_bwd(x, g, ctx.lo, ctx.hi), None, None


def clip(input):
    """RA_HVED.py:749-753: clamp(input, -50, 50), applied to every modality's logvar while the experts are concatenated
    (RA_HVED.py:580).  Stand-alone kernel with torch.clamp's gradient mask; ``ops.poe_fwd_levels(..., clip=(-50, 50))``
    is the fused form (the clamp happens as the PoE kernel loads the slabs)."""
    _require_device(input)
    out = _ClipFunction.apply(input.float().contiguous(), -50.0, 50.0)
    return out.to(input.dtype)


class ZeroLayerF(torch.autograd.Function):
    """buildingblocks.py:308-323 (call sites RA_HVED.py:559, U_Hemis.py:42, buildingblocks.py:879-881):
    ``ZeroLayerF.apply(x, alpha)`` returns a copy of x with the rows selected by the (B,) boolean ``alpha`` zeroed; the
    backward zeroes the same rows of the gradient."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, x, alpha):
        _require_device(x)
        ctx.alpha = alpha
        ctx.dtype = x.dtype
        return ops.zero_rows(x, alpha).to(x.dtype)

    @staticmethod
    @_lib.on_device
    def backward(ctx, grad_output):
        return ops.zero_rows(grad_output, ctx.alpha).to(ctx.dtype), None


class _DiceFunction(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, p, t, eps):
        sums = ops.dice_sums(p, t)
        ctx.save_for_backward(p, t, sums)
        ctx.eps = eps
        return 2.0 * sums[:, 0] / (sums[:, 1] + sums[:, 2]).clamp(min=eps)

    @staticmethod
    @_lib.on_device
    def backward(ctx, g):
        p, t, sums = ctx.saved_tensors
        return ops.dice_bwd(p, t, sums, g.contiguous(), ctx.eps), None, None


def compute_per_channel_dice(input, target, epsilon=1e-6, weight=None):
    """loss.py:257-285: Dice coefficient per channel of (N, C, *spatial) probabilities and targets,
    2 sum(p t) / clamp(sum p^2 + sum t^2, eps), with the optional per-channel weight on the intersection.  One fused pass
    (and one for the gradient w.r.t. ``input``; the target is a label map and receives none)."""
    assert input.size() == target.size(), "'input' and 'target' must have the same shape"
    _require_device(input)
    dice = _DiceFunction.apply(input.float().contiguous(), target.float().contiguous(), float(epsilon))
    if weight is not None:
        dice = dice * weight.reshape(-1).to(dice)
    return dice.to(input.dtype) if input.dtype in (torch.float16, torch.bfloat16) else dice


class DiceLoss(nn.Module):
    """Mirror of loss.DiceLoss (loss.py:188-209): 1 - mean over channels of the Dice coefficient (no normalisation of the
    input: the model ends in a sigmoid, RA_HVED.py:483-484)."""

    def __init__(self, weight=None):
        super().__init__()
        self.weight = weight

    def forward(self, input, target):
        return 1.0 - torch.mean(compute_per_channel_dice(input, target, weight=self.weight))


class _KLDFunction(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, mu5, logvar5, subsets):
        _, _, _, kld = ops.poe_fwd(mu5, logvar5, subsets, want_kld=True)
        n = mu5[0].numel()
        ctx.save_for_backward(mu5, logvar5)
        ctx.subsets, ctx.scale = subsets, 0.5 / (n * len(subsets))
        return kld.sum() * ctx.scale

    @staticmethod
    @_lib.on_device
    def backward(ctx, g):
        mu5, logvar5 = ctx.saved_tensors
        d_mu, d_lv = ops.poe_bwd(mu5, logvar5, ctx.subsets, kld_scale=[ctx.scale] * len(ctx.subsets))
        return d_mu * g, d_lv * g, None


def compute_KLD(mu_list, logvar_list, subset_index_list=[14], choices=[0, 1, 2, 3]):
    """loss.py:85-115: mean over the requested subsets of KL(PoE posterior || prior).  Inputs are the (B,5,...)
    tensors the model returns (RA_HVED.py:582-583); fusion, KL and the reduction run in one launch."""
    mu5 = mu_list.transpose(1, 0).float().contiguous()
    lv5 = logvar_list.transpose(1, 0).float().contiguous()
    subsets = [SUBSETS_MODALITIES[i] for i in range(len(SUBSETS_MODALITIES)) if i in subset_index_list]
    return _KLDFunction.apply(mu5, lv5, subsets)


# ----------------------------------------------------------------------------- conv path: normalisation + LeakyReLU (K6)
def _norm_forward(ctx, x, gamma, beta, mode, eps, slope, mean_in, rstd_in):
    x = x if x.is_contiguous() else x.contiguous()
    gamma = ops._f32c(gamma) if gamma is not None else None
    beta = ops._f32c(beta) if beta is not None else None
    plan = ops.norm_plan(x, mode, eps, slope)
    y, stats = ops.norm_act_fwd_raw(x, gamma, beta, plan, mean_in, rstd_in)
    ctx.plan = plan
    if stats is None:
        ctx.save_for_backward(x, gamma, beta, mean_in, rstd_in)
    else:
        ctx.save_for_backward(x, gamma, beta, stats)
    return y, stats


def _norm_backward(ctx, dy):
    saved = ctx.saved_tensors
    x, gamma, beta = saved[0], saved[1], saved[2]
    if dy.dtype != x.dtype:
        dy = dy.to(x.dtype)
    if not dy.is_contiguous():
        dy = dy.contiguous()
    want = gamma is not None and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
    if len(saved) == 4:
        dx, dgb = ops.norm_act_bwd_raw(x, dy, gamma, beta, ctx.plan, stats=saved[3], want_param_grads=want)
    else:
        dx, dgb = ops.norm_act_bwd_raw(x, dy, gamma, beta, ctx.plan, mean=saved[3], rstd=saved[4], want_param_grads=want)
    if want:
        return dx, dgb[0].to(gamma.dtype), (dgb[1].to(beta.dtype) if beta is not None else None)
    return dx, None, None


class _NormActFunction(torch.autograd.Function):
    """InstanceNorm3d / BatchNorm3d with the LeakyReLU that follows fused in (csrc/norm_act.cu).  Saves x and ONE small buffer
    (mean | rstd | the kernel's scratch) only: the backward recomputes the activation mask."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, x, gamma, beta, mode, eps, slope, mean_in, rstd_in):
        return _norm_forward(ctx, x, gamma, beta, mode, eps, slope, mean_in, rstd_in)[0]

    @staticmethod
    @once_differentiable
    @_lib.on_device
    def backward(ctx, dy):
        return (*_norm_backward(ctx, dy), None, None, None, None, None)


class _NormActStatsFunction(torch.autograd.Function):
    """The same, also returning the batch statistics (mean, rstd) that train-mode BatchNorm folds into its running statistics."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, x, gamma, beta, mode, eps, slope):
        y, stats = _norm_forward(ctx, x, gamma, beta, mode, eps, slope, None, None)
        g = ctx.plan.groups
        mean, rstd = stats[:g], stats[g:2 * g]
        ctx.mark_non_differentiable(mean, rstd)
        return y, mean, rstd

    @staticmethod
    @once_differentiable
    @_lib.on_device
    def backward(ctx, dy, _dmean, _drstd):
        return (*_norm_backward(ctx, dy), None, None, None)


def _ncs(x, num_features):
    """(N, C, *spatial) view of a batched or unbatched input (nn.InstanceNorm3d accepts (C, D, H, W) too)."""
    if x.dim() == 4 and x.shape[0] == num_features:
        return x.unsqueeze(0), True
    return x, False


def instance_norm_act(x, weight=None, bias=None, eps: float = 1e-5, slope: float = 1.0):
    """F.instance_norm(x, weight=weight, bias=bias, eps=eps) followed by F.leaky_relu(., slope) (slope = 1: none) in one pass
    pair over x: statistics per (sample, channel) plane, biased variance."""
    _require_device(x)
    return _NormActFunction.apply(x, weight, bias, ops.NORM_INSTANCE, float(eps), float(slope), None, None)


def batch_norm_act(x, weight, bias, running_mean, running_var, training: bool, momentum, eps: float = 1e-5, slope: float = 1.0):
    """F.batch_norm (+ LeakyReLU).  Training: batch statistics over (N, *spatial), running statistics updated in place with
    ``momentum`` and the unbiased variance; eval: the running statistics."""
    _require_device(x)
    if training:
        y, mean, rstd = _NormActStatsFunction.apply(x, weight, bias, ops.NORM_BATCH, float(eps), float(slope))
        if running_mean is not None and momentum:
            with torch.no_grad():
                n = x.numel() // x.shape[1]
                var = rstd.double().pow(-2).sub_(eps).clamp_(min=0).mul_(n / max(n - 1, 1))
                running_mean.mul_(1 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
                running_var.mul_(1 - momentum).add_(var.to(running_var.dtype), alpha=momentum)
        return y
    rstd = torch.rsqrt(running_var.float() + eps)
    return _NormActFunction.apply(x, weight, bias, ops.NORM_FROZEN, float(eps), float(slope), running_mean.float().contiguous(), rstd)


class InstanceNorm3d(nn.InstanceNorm3d):
    """nn.InstanceNorm3d (the 'i' of SingleConv's order string, buildingblocks.py:430-431; BasicConv.norm, buildingblocks.py:20)
    on the fused kernel.  ``fused_slope`` is the negative slope of a LeakyReLU folded in by ``patch_model`` (None = none)."""
    fused_slope = None

    def forward(self, input):
        return instance_norm_forward(self, input)


def instance_norm_forward(mod, input):
    if mod.track_running_stats:
        raise NotImplementedError("xlstm_hved_b200.InstanceNorm3d: track_running_stats=True is not supported (the reference never sets it)")
    x, unbatched = _ncs(input, mod.num_features)
    slope = getattr(mod, "fused_slope", None)
    y = instance_norm_act(x, mod.weight, mod.bias, mod.eps, 1.0 if slope is None else slope)
    return y.squeeze(0) if unbatched else y


class BatchNorm3d(nn.BatchNorm3d):
    """nn.BatchNorm3d (modules/DuSFE.py:17-36, 108-110, 187) on the fused kernel."""
    fused_slope = None

    def forward(self, input):
        return batch_norm_forward(self, input)


def batch_norm_forward(mod, input):
    """torch.nn.modules.batchnorm._BatchNorm.forward's bookkeeping (momentum None = cumulative average) around batch_norm_act."""
    if input.dim() != 5:
        raise ValueError(f"expected 5D input (got {input.dim()}D input)")
    momentum = 0.0 if mod.momentum is None else mod.momentum
    if mod.training and mod.track_running_stats and mod.num_batches_tracked is not None:
        mod.num_batches_tracked.add_(1)
        if mod.momentum is None:
            momentum = 1.0 / float(mod.num_batches_tracked)
    training = mod.training or (mod.running_mean is None and mod.running_var is None)
    use_running = (not mod.training) or mod.track_running_stats
    slope = getattr(mod, "fused_slope", None)
    return batch_norm_act(input, mod.weight, mod.bias, mod.running_mean if use_running else None,
                          mod.running_var if use_running else None, training, momentum, mod.eps, 1.0 if slope is None else slope)


class FusedAwayLeakyReLU(nn.LeakyReLU):
    """A LeakyReLU whose work happens inside the preceding normalisation kernel (``patch_model`` sets that layer's
    ``fused_slope``): the forward hands its input through."""

    def forward(self, input):
        return input


# ----------------------------------------------------------------------------- conv path: AttenModule2's spatial gate (K7)
class _Gate7Function(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, x, w, b):
        gate = ops.gate7_fwd(x, w, b)
        ctx.save_for_backward(x, w, gate)
        return gate

    @staticmethod
    @once_differentiable
    @_lib.on_device
    def backward(ctx, dgate):
        x, w, gate = ctx.saved_tensors
        dx, dw, db = ops.gate7_bwd(x, w, gate, dgate, want_dx=ctx.needs_input_grad[0],
                                   want_dw=ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        return dx, dw, db


def _conv_out_dtype(x):
    """What a PyTorch convolution would return for this input: the autocast type inside an autocast region (train.py:207),
    the input's type otherwise."""
    if x.is_cuda and torch.is_autocast_enabled("cuda"):
        return torch.get_autocast_dtype("cuda")
    return x.dtype


def gate_convs_supported(dw, pw) -> bool:
    """A depthwise 7^3 convolution (stride 1, zero padding 3, any channel expansion) followed by a 1x1x1 convolution to one channel."""
    return (isinstance(dw, nn.Conv3d) and isinstance(pw, nn.Conv3d) and dw.kernel_size == (7, 7, 7) and dw.stride == (1, 1, 1)
            and dw.padding == (3, 3, 3) and dw.dilation == (1, 1, 1) and dw.groups == dw.in_channels and dw.padding_mode == "zeros"
            and dw.out_channels % dw.in_channels == 0 and pw.kernel_size == (1, 1, 1) and pw.stride == (1, 1, 1)
            and pw.padding == (0, 0, 0) and pw.groups == 1 and pw.in_channels == dw.out_channels and pw.out_channels == 1)


def spatial_gate(x, dw, pw):
    """sigmoid(pw(dw(x))) of AttenModule2 (buildingblocks.py:283-285, 294-296) as ONE dense G -> 1 convolution + sigmoid kernel:
    the two linear layers are composed here (two small torch ops; autograd carries the gradient back to both layers' parameters)."""
    _require_device(x)
    if not gate_convs_supported(dw, pw):
        raise NotImplementedError("xlstm_hved_b200.spatial_gate: expects a depthwise 7x7x7 conv followed by a 1x1x1 conv to one channel")
    G, e = dw.in_channels, dw.out_channels // dw.in_channels
    w2 = pw.weight.reshape(G, e, 1).float()
    w = (dw.weight.reshape(G, e, 343).float() * w2).sum(1)
    b = None
    if dw.bias is not None:
        b = (pw.weight.reshape(-1).float() * dw.bias.float()).sum().reshape(1)
    if pw.bias is not None:
        b = pw.bias.float().reshape(1) if b is None else b + pw.bias.float().reshape(1)
    return _Gate7Function.apply(x.float(), w, b).to(_conv_out_dtype(x))


def atten_module2_forward(mod, seg_x, enc_x, recon_x=None):
    """AttenModule2.forward (buildingblocks.py:277-301) with both spatial gates on the fused kernel; everything else as written
    there (ChannelPool, the two scalings, the concatenation)."""
    spa_comp = mod.compress(seg_x)
    enc_spa = torch.cat([spa_comp, mod.compress(enc_x)], 1)
    enc_scale = spatial_gate(enc_spa, mod.enc_spatial, mod.enc_spatial2)
    s_enc_x = enc_x + enc_x * enc_scale
    if recon_x is not None:
        raise NameError("name 'comp_x' is not defined")        # buildingblocks.py:289-290: the reference fails here too
    seg_scale = spatial_gate(spa_comp, mod.seg_spatial, mod.seg_spatial2)
    scaled_seg_x = seg_x * (1 + seg_scale)
    return torch.cat([scaled_seg_x, s_enc_x], 1)


# ----------------------------------------------------------------------------- conv path: depthwise 3x3x3 convolution (K8)
class _DwConv3Function(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, x, w, b):
        y = ops.dwconv3_fwd(x, w, b)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    @once_differentiable
    @_lib.on_device
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx, dw, db = ops.dwconv3_bwd(x, w, dy, want_dx=ctx.needs_input_grad[0], want_dw=ctx.needs_input_grad[1],
                                     want_db=ctx.has_bias and ctx.needs_input_grad[2])
        return dx, dw, db


def dwconv3_supported(conv) -> bool:
    """nn.Conv3d(C, C, 3, stride 1, zero padding 1, groups = C): the conv of BasicConv(C, C, 3, padding=1, groups=C), RA_HVED.py:406."""
    return (isinstance(conv, nn.Conv3d) and conv.kernel_size == (3, 3, 3) and conv.stride == (1, 1, 1) and conv.padding == (1, 1, 1)
            and conv.dilation == (1, 1, 1) and conv.padding_mode == "zeros" and conv.groups == conv.in_channels == conv.out_channels)


def depthwise_conv3_forward(conv, x):
    """nn.Conv3d.forward of a depthwise 3x3x3 layer on the fused kernels (forward, input gradient, weight gradient)."""
    _require_device(x)
    unbatched = x.dim() == 4
    xb = x.unsqueeze(0) if unbatched else x
    y = _DwConv3Function.apply(xb.float(), conv.weight.float(), conv.bias.float() if conv.bias is not None else None)
    y = y.to(_conv_out_dtype(x))
    return y.squeeze(0) if unbatched else y


# ----------------------------------------------------------------------------- conv path: 1x1x1 convolution (K9)
class _PwConvFunction(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, x, w, b):
        y = ops.pwconv_fwd(x, w, b)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    @once_differentiable
    @_lib.on_device
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx, dw, db = ops.pwconv_bwd(x, w, dy, want_dx=ctx.needs_input_grad[0], want_dw=ctx.needs_input_grad[1],
                                    want_db=ctx.has_bias and ctx.needs_input_grad[2])
        return dx, (dw.to(w.dtype) if dw is not None else None), db


def pwconv_supported(conv) -> bool:
    """nn.Conv3d(Cin, Cout, 1) with stride 1, no padding, one group, at most 32 channels on either side."""
    return (isinstance(conv, nn.Conv3d) and conv.kernel_size == (1, 1, 1) and conv.stride == (1, 1, 1) and conv.padding == (0, 0, 0)
            and conv.dilation == (1, 1, 1) and conv.groups == 1 and conv.in_channels <= 32 and conv.out_channels <= 32)


def pointwise_conv_forward(conv, x):
    """nn.Conv3d.forward of a 1x1x1 layer on the fused kernels.  Inside an autocast region the input is taken (and the output
    returned) in the autocast type, as PyTorch's convolution would (train.py:207); weights, bias and accumulation stay fp32."""
    _require_device(x)
    unbatched = x.dim() == 4
    xb = x.unsqueeze(0) if unbatched else x
    xb = xb.to(_conv_out_dtype(xb))
    if xb.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        xb = xb.float()
    y = _PwConvFunction.apply(xb, conv.weight.float(), conv.bias.float() if conv.bias is not None else None)
    return y.squeeze(0) if unbatched else y


# ----------------------------------------------------------------------------- conv path: dense 3x3x3 convolution, few channels (K10)
class _Conv3Function(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, x, w, b):
        y = ops.conv3_fwd(x, w, b)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    @once_differentiable
    @_lib.on_device
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx, dw, db = ops.conv3_bwd(x, w, dy, want_dx=ctx.needs_input_grad[0], want_dw=ctx.needs_input_grad[1],
                                   want_db=ctx.has_bias and ctx.needs_input_grad[2])
        return dx, (dw.to(w.dtype) if dw is not None else None), db


def conv3_supported(conv) -> bool:
    """nn.Conv3d(Cin, Cout, 3, stride 1, zero padding 1), one group, at most 64 channels on either side."""
    return (isinstance(conv, nn.Conv3d) and conv.kernel_size == (3, 3, 3) and conv.stride == (1, 1, 1) and conv.padding == (1, 1, 1)
            and conv.dilation == (1, 1, 1) and conv.padding_mode == "zeros" and conv.groups == 1 and conv.in_channels <= 64
            and conv.out_channels <= 64)


def dense_conv3_forward(conv, x):
    """nn.Conv3d.forward of a small-channel 3x3x3 layer on the direct-convolution kernels; autocast behaviour as in
    pointwise_conv_forward."""
    _require_device(x)
    unbatched = x.dim() == 4
    xb = x.unsqueeze(0) if unbatched else x
    xb = xb.to(_conv_out_dtype(xb))
    if xb.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        xb = xb.float()
    y = _Conv3Function.apply(xb, conv.weight.float(), conv.bias.float() if conv.bias is not None else None)
    return y.squeeze(0) if unbatched else y
