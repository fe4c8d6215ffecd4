"""coskad_b200 -- B200-native (sm_100a) implementation of COSKAD's anomaly-scoring hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed plumbing); every hot op runs
in hand-written CUDA kernels behind the C ABI in include/coskad_b200.h (libcoskad_b200.so, loaded
with ctypes by coskad_b200._lib).  There is no CPU / eager fallback.
"""
from . import _lib
from ._lib import (CoskadError, Context, SCORE_NONE, SCORE_POINCARE, SCORE_POINCARE_NOPROJ, SCORE_EUCLID,
                   SCORE_COSINE, SCORE_POINCARE_HM)

__all__ = ['_lib', 'CoskadError', 'Context', 'SCORE_NONE', 'SCORE_POINCARE', 'SCORE_POINCARE_NOPROJ',
           'SCORE_EUCLID', 'SCORE_COSINE', 'SCORE_POINCARE_HM']
__version__ = '0.1.0'
