"""phones_las_b200 -- B200-native (sm_100a) implementation of the phones-las hot path:
MFCC/MFE front-end -> pyramidal BiLSTM listener -> attention decoder (greedy / teacher-forced).

Host side is Python over PyTorch tensors; all arithmetic runs in hand-written CUDA reached
through the C-ABI in include/plas.h (libplas.so, loaded with ctypes by ``_lib``).  There is
no CPU fallback: ops raise if the library is missing.
"""
from .hparams import (create_hparams, get_default_hparams, feature_args, baseline_config,  # noqa: F401
                      SAMPLE_RATE, SOS_ID, EOS_ID, UNK_ID)

__all__ = ["create_hparams", "get_default_hparams", "feature_args", "baseline_config",
           "SAMPLE_RATE", "SOS_ID", "EOS_ID", "UNK_ID"]
