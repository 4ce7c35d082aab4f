"""SGD with one fused multi-tensor kernel per step.

The reference uses ``torch.optim.SGD(model.parameters(), 1e-4)`` (ref: ddp_example.py:62,90,92),
i.e. ``p -= lr·g`` through a foreach kernel, with ``zero_grad()`` dropping the gradients so
autograd re-allocates them every backward.  Ours:

* the update for *all* parameters is one sm_100a kernel launch (``ops.sgd_step``) over a pointer
  table — momentum, dampening, Nesterov and weight decay included — so it is graph-capturable
  and launch-count-free as the model grows;
* ``lr`` lives in a device scalar when ``capturable=True`` so schedulers work under CUDA graphs;
* ``zero_grad(set_to_none=False)`` keeps gradients as views of the DDP bucket (no re-allocation,
  no copy into the bucket next step). ``set_to_none=True`` (the reference's default behaviour)
  still works — the reducer re-homes gradients on the next backward.

Bookkeeping (param_groups / state_dict) comes from ``torch.optim.Optimizer``.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch


class SGD(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, momentum: float = 0.0, dampening: float = 0.0,
                 weight_decay: float = 0.0, nesterov: bool = False, maximize: bool = False,
                 capturable: bool = False, fused: Optional[bool] = None):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if momentum < 0.0:
            raise ValueError(f"Invalid momentum value: {momentum}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                        nesterov=nesterov, maximize=maximize, capturable=capturable, fused=fused)
        super().__init__(params, defaults)

    def zero_grad(self, set_to_none: bool = True) -> None:
        if set_to_none:
            return super().zero_grad(set_to_none=True)
        grads = [p.grad for g in self.param_groups for p in g["params"] if p.grad is not None]
        if grads:
            torch._foreach_zero_(grads)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        from .. import ops

        for group in self.param_groups:
            params, grads, bufs = [], [], []
            momentum = group["momentum"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("SGD does not support sparse gradients")
                params.append(p)
                grads.append(p.grad)
                if momentum != 0:
                    st = self.state[p]
                    if "momentum_buffer" not in st:
                        st["momentum_buffer"] = None
                    bufs.append(st)
            if not params:
                continue
            use_fused = group["fused"]
            if use_fused is None:
                use_fused = params[0].is_cuda and ops.native_available() and all(
                    p.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous() and g.dtype == torch.float32
                    for p, g in zip(params, grads))
            if use_fused:
                first = False
                mbufs = None
                if momentum != 0:
                    first = any(st["momentum_buffer"] is None for st in bufs)
                    for st, p in zip(bufs, params):
                        if st["momentum_buffer"] is None:
                            st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    mbufs = [st["momentum_buffer"] for st in bufs]
                ops.sgd_step(params, grads, mbufs, lr=group["lr"], momentum=momentum, dampening=group["dampening"],
                             weight_decay=group["weight_decay"], nesterov=group["nesterov"],
                             maximize=group["maximize"], first_step=first)
                continue
            # reference math through foreach ops (CPU / exotic dtypes)
            gs = [(-g if group["maximize"] else g) for g in grads] if group["maximize"] else list(grads)
            if group["weight_decay"] != 0:
                gs = torch._foreach_add(gs, params, alpha=group["weight_decay"])
            if momentum != 0:
                new = []
                for st, g in zip(bufs, gs):
                    if st["momentum_buffer"] is None:
                        st["momentum_buffer"] = torch.clone(g).detach()
                    else:
                        st["momentum_buffer"].mul_(momentum).add_(g, alpha=1 - group["dampening"])
                    new.append(st["momentum_buffer"])
                gs = torch._foreach_add(gs, new, alpha=momentum) if group["nesterov"] else new
            torch._foreach_add_(params, gs, alpha=-group["lr"])
        return loss
