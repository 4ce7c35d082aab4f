"""Gradient communication hooks for :class:`DistributedDataParallel`.

Interface parity: ``ddp.register_comm_hook(state, hook)`` with
``hook(state, bucket: GradBucket) -> Future[Tensor]`` (torch c10d comm.hpp:20-118,
ddp_comm_hooks/default_hooks.py:35-175).  The bucket a hook receives holds *undivided* local
gradients (observed behaviour of the reference stack); the hook owns the
averaging.  The reference script itself installs no hook (ref: ddp_example.py:64) — the built-in
path is the fused reduce kernel — so these are the optional surface.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from .. import distributed as dist


class Future:
    """Minimal future: a value that becomes available after a collective's ``wait()``."""

    def __init__(self, work=None, value: Optional[torch.Tensor] = None):
        self._work = work
        self._value = value
        self._callbacks: List[Callable] = []
        self._done = work is None

    def then(self, fn: Callable[["Future"], torch.Tensor]) -> "Future":
        self._callbacks.append(fn)
        return self

    def wait(self) -> torch.Tensor:
        if not self._done:
            self._work.wait()
            self._done = True
        for fn in self._callbacks:
            self._value = fn(self)
        self._callbacks = []
        return self._value

    def value(self) -> torch.Tensor:
        return self.wait()

    def done(self) -> bool:
        return self._done or self._work.is_completed()


def _group(state):
    return state if isinstance(state, dist.ProcessGroup) else dist.get_default_group()


def allreduce_hook(state, bucket) -> Future:
    """Average gradients: sum-allreduce with the 1/world scale folded into the collective."""
    g = _group(state)
    buf = bucket.buffer()
    return Future(g.comm.allreduce(buf, dist.ReduceOp.SUM, 1.0 / g.size()), buf)


def _compress_hook(dtype):
    def hook(state, bucket) -> Future:
        g = _group(state)
        buf = bucket.buffer()
        wire = buf.to(dtype).div_(g.size())
        fut = Future(g.comm.allreduce(wire, dist.ReduceOp.SUM, 1.0), buf)

        def decompress(_f):
            buf.copy_(wire)
            return buf

        return fut.then(decompress)

    return hook


fp16_compress_hook = _compress_hook(torch.float16)
bf16_compress_hook = _compress_hook(torch.bfloat16)


def noop_hook(state, bucket) -> Future:
    """Skips communication (for measuring pure compute); gradients stay local."""
    return Future(None, bucket.buffer())


def fp16_compress_wrapper(hook):
    def wrapped(state, bucket):
        buf = bucket.buffer()
        half = buf.to(torch.float16)
        bucket.set_buffer(half)
        fut = hook(state, bucket)

        def back(f):
            buf.copy_(f._value if isinstance(f, Future) else f)
            return buf

        return fut.then(back)

    return wrapped


class PostLocalSGDState:
    """State for :func:`post_localSGD_hook`: global allreduce for the first ``start_localSGD_iter``
    iterations, then no gradient communication (parameters are averaged periodically by the user)."""

    def __init__(self, process_group=None, start_localSGD_iter: int = 100):
        self.process_group = process_group
        self.start_localSGD_iter = start_localSGD_iter
        self.iter = 0

    def maybe_increase_iter(self, bucket):
        if bucket.is_last():
            self.iter += 1


def post_localSGD_hook(state: PostLocalSGDState, bucket) -> Future:
    if state.iter < state.start_localSGD_iter:
        fut = allreduce_hook(state.process_group, bucket)
    else:
        fut = noop_hook(None, bucket)
    state.maybe_increase_iter(bucket)
    return fut
