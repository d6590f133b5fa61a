"""xlstm_hved_b200 -- B200-native (sm_100a) ViL-mLSTM + S-MVAE hot path of XLSTM-HVED.

CUDA only: importing is cheap, but every op needs libxhved.so (python -m xlstm_hved_b200.build)
and a CUDA device; there is no CPU / PyTorch-eager fallback.
"""
from . import dist, driver, modules, ops  # noqa: F401
from .driver import GraphedSubsetsForward, all_subsets_forward  # noqa: F401
from .graph import GraphedStep  # noqa: F401
from .modules import (BatchNorm3d, DiceLoss, InstanceNorm3d, ProductOfExperts, ProductOfExperts2, SequenceTraversal, ViLBlock, ViLLayer, ViLLayer3D,  # noqa: F401
                      batch_norm_act, clip, compute_KLD, depthwise_conv3_forward, instance_norm_act, parallel_stabilized_simple, reparametrize,
                      spatial_gate)
from .patch import patch_model, unpatch_model  # noqa: F401
