"""sm_100a operator library (Python face).  Every function here launches hand-written CUDA
kernels from ``_C`` (csrc/cuda/*.cu); there is no eager fallback on a GPU: if the native runtime
is missing on a CUDA machine the ops raise instead of silently running ATen kernels."""
from __future__ import annotations

import torch

from .. import _C

_HAS_NATIVE = hasattr(_C, "ops_ready")


def native_available() -> bool:
    """True when the compiled sm_100a kernels are present *and* a CUDA device is usable."""
    return _HAS_NATIVE and torch.cuda.is_available()


def require_native(what: str) -> None:
    if not _HAS_NATIVE:
        raise RuntimeError(f"{what}: the native sm_100a runtime (_C.so with CUDA kernels) is missing; "
                           "run `python -m pytorch_distributed_train_b200._build`")
    if not torch.cuda.is_available():
        raise RuntimeError(f"{what}: no CUDA device available")


if _HAS_NATIVE:
    from .functional import (bn_apply, bn_backward_apply, bn_backward_reduce, bn_finalize, bn_local_stats,  # noqa: F401
                             conv_bn_relu_pool, conv2d, cross_entropy, linear, sgd_step)
else:  # CPU-only build of the extension: keep the names importable, fail loudly on use
    def _missing(name):
        def f(*a, **k):
            require_native(name)
        f.__name__ = name
        return f

    for _n in ("bn_apply", "bn_backward_apply", "bn_backward_reduce", "bn_finalize", "bn_local_stats", "conv_bn_relu_pool",
               "conv2d", "cross_entropy", "linear", "sgd_step"):
        globals()[_n] = _missing(_n)
