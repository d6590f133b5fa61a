"""xlstm_hved_b200 -- B200-native (sm_100a) ViL-mLSTM + S-MVAE hot path of XLSTM-HVED.

CUDA only: importing is cheap, but every op needs libxhved.so (python -m xlstm_hved_b200.build)
and a CUDA device; there is no CPU / PyTorch-eager fallback.
"""
from . import ops  # noqa: F401

__all__ = ["ops"]
