"""CPU-side checks: the C-ABI library loads and exports every symbol include/xhved.h declares, the mirror
modules keep the reference's state_dict keys, patching binds and restores, and nothing computes without CUDA."""
import os
import re

import pytest
import torch

from conftest import ROOT, load_golden


def test_library_exports_every_declared_symbol():
    from xlstm_hved_b200 import _lib
    header = open(os.path.join(ROOT, "include", "xhved.h")).read()
    declared = set(re.findall(r"\b(xhved_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load_library()                       # raises if the .so is missing or lacks a symbol
    assert lib.xhved_version() >= 100
    names = [lib.xhved_profile_kernel_name(i).decode() for i in range(lib.xhved_profile_kernel_count())]
    assert "mlstm_chunk_grad" in names and "poe_fwd" in names


def test_mirror_modules_keep_reference_state_dict_keys():
    import xlstm_hved_b200 as xh
    c = load_golden("vil_wrapper.pt")
    wrap = xh.ViLLayer3D(dim=32)
    assert set(wrap.state_dict().keys()) == set(c["state_dict"].keys())
    wrap.load_state_dict(c["state_dict"], strict=True)
    for k, v in wrap.state_dict().items():
        assert v.shape == c["state_dict"][k].shape, k
    blk = load_golden("vil_block.pt")["dim64_s140_rev"]
    b = xh.ViLBlock(dim=64, direction=xh.SequenceTraversal.ROWWISE_FROM_BOT_RIGHT)
    b.load_state_dict(blk["state_dict"], strict=True)


def test_no_cpu_fallback():
    import xlstm_hved_b200 as xh
    with pytest.raises(RuntimeError):
        xh.parallel_stabilized_simple(*[torch.zeros(1, 1, 4, 16) for _ in range(3)], torch.zeros(1, 1, 4, 1), torch.zeros(1, 1, 4, 1))
    with pytest.raises(RuntimeError):
        xh.ViLLayer3D(dim=32)(torch.zeros(1, 32, 2, 2, 2))
    with pytest.raises(RuntimeError):
        xh.ProductOfExperts()(torch.zeros(5, 1, 1, 2, 2, 2), torch.zeros(5, 1, 1, 2, 2, 2), (0, 1))
    # the conv-path ops and the graphed driver refuse as well
    with pytest.raises(RuntimeError):
        xh.instance_norm_act(torch.zeros(1, 2, 4, 4, 4))
    with pytest.raises(RuntimeError):
        xh.spatial_gate(torch.zeros(1, 2, 4, 4, 4), torch.nn.Conv3d(2, 8, 7, padding=3, groups=2), torch.nn.Conv3d(8, 1, 1))
    with pytest.raises(RuntimeError):
        xh.ops.conv3_fwd(torch.zeros(1, 2, 4, 4, 4), torch.zeros(2, 2, 3, 3, 3))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            xh.GraphedSubsetsForward(torch.nn.Identity().eval())


def test_patch_and_unpatch_reference_model():
    from oracle import ref_loader
    if ref_loader.find_reference() is None:
        pytest.skip("reference tree not present on this machine")
    import io
    import sys
    import types
    import xlstm_hved_b200 as xh
    model = ref_loader.build_model()
    ns = ref_loader.load_reference()
    # train.py:23 / Pretrain.py:23 bind the loss function by name at import time: the patch must reach such bindings too
    fake_train = types.ModuleType("xhved_fake_train")
    fake_train.compute_KLD = ns.loss.compute_KLD
    sys.modules["xhved_fake_train"] = fake_train
    try:
        stock_cls = type(model.mViL.vil)
        counts = xh.patch_model(model)
        assert counts["ViLBlock"] == 1 and counts["ProductOfExperts"] == 1 and counts["ProductOfExperts2"] == 1
        assert "xhved_fake_train.compute_KLD" in counts["rebound"] and "loss.compute_KLD" in counts["rebound"]
        assert "RA_HVED.reparametrize" in counts["rebound"] and "RA_HVED.clip" in counts["rebound"]
        assert "buildingblocks.ZeroLayerF" in counts["rebound"] and "RA_HVED.ZeroLayerF" in counts["rebound"]
        assert fake_train.compute_KLD is xh.compute_KLD and ns.RA_HVED.ZeroLayerF is xh.modules.ZeroLayerF
        # no instance-level state: the override lives on a subclass, so DataParallel replicas (which copy __dict__,
        # train.py:148-151) resolve `self` to the replica
        vil = model.mViL.vil
        assert "forward" not in vil.__dict__ and type(vil) is not stock_cls and isinstance(vil, stock_cls)
        replica = vil._replicate_for_data_parallel()
        assert type(replica) is type(vil) and replica.forward.__self__ is replica
        # train.py:370-397 pickles the whole model: a patched model must save, and load back as the stock classes
        buf = io.BytesIO()
        torch.save({"model": model}, buf)
        buf.seek(0)
        back = torch.load(buf, weights_only=False)["model"]
        assert type(back.mViL.vil) is stock_cls and type(back.experts).__name__ == "ProductOfExperts"
        assert not hasattr(type(back.experts), "_xhved_base")
        keys = set(model.state_dict().keys())
        assert set(back.state_dict().keys()) == keys
        xh.unpatch_model(model)
        assert type(model.mViL.vil) is stock_cls and set(model.state_dict().keys()) == keys
        assert ns.RA_HVED.reparametrize.__module__ == "RA_HVED" and ns.RA_HVED.clip.__module__ == "RA_HVED"
        assert fake_train.compute_KLD is ns.loss.compute_KLD and fake_train.compute_KLD.__module__ == "loss"
        assert ns.RA_HVED.ZeroLayerF is ns.buildingblocks.ZeroLayerF and ns.RA_HVED.ZeroLayerF.__module__ == "buildingblocks"
    finally:
        del sys.modules["xhved_fake_train"]
    with torch.no_grad():
        seg, _ = model.eval()(torch.rand(1, 4, 32, 32, 32), [14], valid=True)   # stock path still runs
    assert seg.shape == (1, 3, 32, 32, 32)


def test_patched_classes_without_a_reference_tree():
    """The same mechanics on a stand-in class (the GPU box has no reference tree): class swap, replica binding, pickling
    as the original class."""
    import io
    import xlstm_hved_b200 as xh
    from xlstm_hved_b200 import patch

    blk = _StandInViLBlock()
    model = torch.nn.Sequential(blk)
    counts = xh.patch_model(model, patch_globals=False)
    assert counts["ViLBlock"] == 1 and type(blk) is not _StandInViLBlock and isinstance(blk, _StandInViLBlock)
    assert xh.patch_model(model, patch_globals=False)["ViLBlock"] == 0          # idempotent
    assert type(blk).forward is patch._vil_forward
    rep = blk._replicate_for_data_parallel()
    assert rep.forward.__self__ is rep
    buf = io.BytesIO()
    torch.save(model, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    assert type(back[0]) is _StandInViLBlock
    xh.unpatch_model(model)
    assert type(blk) is _StandInViLBlock


class _Cell(torch.nn.Module):
    pass


class _Layer(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.mlstm_cell = _Cell()


class ViLBlock(torch.nn.Module):                    # stand-in: patch_model recognises the reference class by name + structure
    def __init__(self):
        super().__init__()
        self.layer = _Layer()
        self.w = torch.nn.Parameter(torch.zeros(3))

    def forward(self, x):
        return x


_StandInViLBlock = ViLBlock


def test_workspace_queries_match_the_host_layer():
    """xhved_*_workspace_query (host-only) against the sizes the Python layer derives from shapes and parameters."""
    import ctypes
    import xlstm_hved_b200 as xh
    from xlstm_hved_b200 import _lib
    lib = _lib.load_library()
    for dim, B, S in ((16, 1, 150), (32, 2, 4096), (64, 3, 6144)):
        w = _lib.VilWorkspaceSizes()
        assert lib.xhved_vil_workspace_query(B, S, dim, ctypes.byref(w)) == 0
        E, nc = 2 * dim, (S + 127) // 128
        dhp = max(E // 4, 16)
        assert (w.cell.nc, w.cell.dhp) == (nc, dhp)
        assert w.cell.tile_bytes == 4 * B * nc * 128 * dhp * 2
        assert w.cell.row_bytes == 4 * B * nc * 128 * 4 and w.cell.chunk_bytes == 4 * B * nc * 4
        assert w.cell.dstate_bytes == 4 * B * nc * dhp * (dhp + 16) * 4 == w.cell.states_bytes
        assert w.token_tile_bytes == B * nc * E * 128 * 2
        blk = xh.ViLBlock(dim, xh.SequenceTraversal.ROWWISE_FROM_TOP_LEFT)
        n_params = sum(p.numel() for p in xh.modules.vil_block_params(blk))
        assert w.grad_replica_stride == (n_params + 31) // 32 * 32
        # the one-call block entry points carve their buffers out of two blobs: at least the sum of the parts
        sv, sc, npg = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        assert lib.xhved_vil_block_workspace(B, S, dim, 32, ctypes.byref(sv), ctypes.byref(sc), ctypes.byref(npg)) == 0
        assert npg.value == n_params
        assert sv.value >= 4 * w.cell.tile_bytes + w.cell.states_bytes + 4 * w.cell.row_bytes + w.cell.dstate_bytes + 3 * w.token_tile_bytes
        assert sc.value >= 32 * w.grad_replica_stride * 4 + 4 * w.cell.tile_bytes + w.cell.states_bytes + 4 * w.token_tile_bytes
    bad = _lib.MlstmWorkspace()
    assert lib.xhved_mlstm_workspace_query(4, 100, 200, ctypes.byref(bad)) == -2      # XHVED_ERR_UNSUPPORTED_DH
    assert lib.xhved_vil_workspace_query(1, 100, 48, ctypes.byref(_lib.VilWorkspaceSizes())) == -4   # XHVED_ERR_UNSUPPORTED_DIM


def test_conv_path_patching_host_logic():
    """patch_model's conv-path kinds (K6 - K10) on CPU: what is taken over and what is left alone, the LeakyReLU folding, idempotence,
    conv_path=False, pickling as the stock classes, unpatch -- and that a patched layer refuses CPU tensors (no fallback)."""
    import io
    import torch.nn as nn
    import xlstm_hved_b200 as xh

    def build():
        return nn.Sequential(
            nn.Sequential(nn.InstanceNorm3d(4), nn.LeakyReLU(0.01, inplace=True), nn.Conv3d(4, 4, 3, padding=1)),      # SingleConv 'ilc'
            nn.Conv3d(4, 4, 3, padding=1, groups=4, bias=False),          # depthwise (K8)
            nn.Conv3d(4, 1, 1),                                            # 1x1x1 (K9)
            nn.Conv3d(4, 2, 3, stride=2, padding=1),                       # stride 2: stays on cuDNN
            nn.Conv3d(4, 4, 3, padding=1, padding_mode="replicate"),       # not zero padding: left alone
            nn.Conv3d(128, 8, 3, padding=1),                               # too many input channels: left alone
            nn.InstanceNorm3d(4, affine=True), nn.ReLU(),                  # norm taken, ReLU not fused
            nn.BatchNorm3d(4), nn.GroupNorm(2, 4))
    m = build()
    counts = xh.patch_model(m)
    assert (counts["InstanceNorm3d"], counts["BatchNorm3d"], counts["fused_LeakyReLU"]) == (2, 1, 1)
    assert (counts["DenseConv3d"], counts["DepthwiseConv3d"], counts["PointwiseConv3d"]) == (1, 1, 1)
    assert m[0][0].fused_slope == pytest.approx(0.01) and getattr(m[6], "fused_slope", None) is None
    for keep in (m[3], m[4], m[5], m[9]):
        assert not hasattr(type(keep), "_xhved_base")
    assert isinstance(m[0][1], nn.LeakyReLU) and type(m[0][1]) is not nn.LeakyReLU
    again = xh.patch_model(m)
    assert all(again[k] == 0 for k in ("InstanceNorm3d", "BatchNorm3d", "DenseConv3d", "fused_LeakyReLU"))      # idempotent
    with pytest.raises(RuntimeError):
        m[0](torch.zeros(1, 4, 4, 4, 4))                                   # CPU tensor: the patched layers have no fallback
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    assert all(type(a) is type(b) for a, b in zip(back.modules(), build().modules()))
    keys = set(m.state_dict().keys())
    xh.unpatch_model(m)
    assert all(type(a) is type(b) for a, b in zip(m.modules(), build().modules())) and set(m.state_dict().keys()) == keys
    assert "fused_slope" not in m[0][0].__dict__
    m2 = build()
    off = xh.patch_model(m2, conv_path=False)
    assert all(off[k] == 0 for k in ("InstanceNorm3d", "BatchNorm3d", "DenseConv3d", "DepthwiseConv3d", "PointwiseConv3d", "fused_LeakyReLU"))


def test_conv_path_patch_counts_on_the_reference_model():
    from oracle import ref_loader
    if ref_loader.find_reference() is None:
        pytest.skip("reference tree not present on this machine")
    import xlstm_hved_b200 as xh
    model = ref_loader.build_model()
    keys = set(model.state_dict().keys())
    counts = xh.patch_model(model, patch_globals=False)
    try:
        assert counts["InstanceNorm3d"] == 80 and counts["fused_LeakyReLU"] == 80 and counts["BatchNorm3d"] == 18
        assert counts["AttenModule2"] == 3 and counts["DepthwiseConv3d"] == 18 and counts["PointwiseConv3d"] == 44
        assert counts["DenseConv3d"] == 62
        assert set(model.state_dict().keys()) == keys
    finally:
        xh.unpatch_model(model)
    with torch.no_grad():
        seg, _ = model.eval()(torch.rand(1, 4, 32, 32, 32), [14], valid=True)   # stock path still runs after unpatch
    assert seg.shape == (1, 3, 32, 32, 32)


def test_shared_activation_module_is_not_folded():
    """One LeakyReLU instance registered under two parents: folding it into the first norm would silently remove the activation
    of the second chain, so neither is folded (the norms still move onto the kernel)."""
    import torch.nn as nn
    import xlstm_hved_b200 as xh
    act = nn.LeakyReLU(0.01)
    m = nn.Sequential(nn.Sequential(nn.InstanceNorm3d(2), act), nn.Sequential(nn.Conv3d(2, 2, 5), act))
    counts = xh.patch_model(m)
    assert counts["InstanceNorm3d"] == 1 and counts["fused_LeakyReLU"] == 0 and type(act) is nn.LeakyReLU
    assert getattr(m[0][0], "fused_slope", None) is None
    xh.unpatch_model(m)
