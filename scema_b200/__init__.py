"""scema_b200 — B200-native implementation of SCEMa's MD-redundancy clustering hot path.

The product is the C-ABI shared library ``scema_b200/libscema_hist.so`` (hand-written CUDA for
sm_100a, declared in ``include/scema_hist.h``) plus the C++ drop-in header and command lines under
``scema_b200/host``. This Python package is a thin ctypes binding over that C ABI, used by the
tests, ``bench.py`` and the multi-GPU driver (``scema_b200.distributed``); PyTorch appears only
as plumbing (device memory, streams, ``torch.distributed``).

There is no CPU fallback: importing works anywhere (so the build can be checked on a machine
without a GPU), but creating a :class:`HistCluster` without a CUDA device raises.
"""
from .binding import (  # noqa: F401
    HistCluster,
    MultiCluster,
    ScemaError,
    PAIRS_DMMA,
    PAIRS_FMA,
    PAIRS_EXACT,
    PAIRS_TC,
    lib,
    lib_path,
    reduce_dir,
)

__all__ = ["HistCluster", "MultiCluster", "ScemaError", "PAIRS_DMMA", "PAIRS_FMA", "PAIRS_EXACT", "PAIRS_TC", "lib", "lib_path",
           "reduce_dir"]
