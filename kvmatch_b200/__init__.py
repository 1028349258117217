"""kvmatch_b200 — B200-native (sm_100a) phase-2 candidate verification and window-mean index pass of KV-match.

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI of include/kvmatch_gpu.h), the ctypes
binding, the host-side mirror of the reference's engine classes, seeded DataGenerator-style inputs
and the multi-GPU offset sharding.  There is no CPU fallback.
"""
from ._lib import KvmError, LIB_PATH  # noqa: F401
from .engine import (GpuSeries, IndexBuilder, MultiGpuSeries, NormQueryEngine, NormQueryEngineDtw, QueryEngine,  # noqa: F401
                     QueryEngineDtw, StatisticInfo, VerifyResult, WU_LIST, rho_from_prompt)
