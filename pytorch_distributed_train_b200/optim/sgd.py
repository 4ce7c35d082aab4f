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
        self._ddp = None            # set by fuse_with_ddp()
        self._fused_active = False
        self._flat_momentum = None
        self._lr_dev = {}           # group index -> (device scalar, last host value) when capturable

    # ---- device-resident learning rate (CUDA-graph friendly schedulers) ------------------------------
    def _lr_tensor(self, gi: int, group, device) -> Optional[torch.Tensor]:
        if not group.get("capturable") or device.type != "cuda":
            return None
        ent = self._lr_dev.get(gi)
        if ent is None:
            ent = [torch.full((1,), float(group["lr"]), dtype=torch.float32, device=device), float(group["lr"])]
            self._lr_dev[gi] = ent
        elif ent[1] != float(group["lr"]) and not torch.cuda.is_current_stream_capturing():
            ent[0].fill_(float(group["lr"]))
            ent[1] = float(group["lr"])
        return ent[0]

    def sync_lr(self) -> None:
        """Push ``param_groups[i]['lr']`` into the device scalars a captured step reads (call between
        graph replays after a scheduler step; a no-op when nothing changed)."""
        for gi, group in enumerate(self.param_groups):
            ent = self._lr_dev.get(gi)
            if ent is not None and ent[1] != float(group["lr"]):
                ent[0].fill_(float(group["lr"]))
                ent[1] = float(group["lr"])

    # ---- single GPU: the update rides on the model's last backward kernel ---------------------------------
    def ride_on_backward(self, model) -> bool:
        """Let the reference ConvNet's last backward kernel apply this optimizer's update (csrc/cuda/fused_convnet.cu: SgdRider): the
        thread that writes a folded gradient element updates the parameter with the value still in its register; the parameters whose
        gradients are complete earlier are updated in the shadow of that kernel's first grid barrier.  ``step()`` then has only
        bookkeeping left.  Used by ``engine.GraphedTrainStep`` when the gradient reduction does not already carry the update (one GPU).

        Same contract as :meth:`fuse_with_ddp`: between ``backward()`` and ``step()`` the parameters are already updated.  Returns
        False (and changes nothing) when the model / optimizer combination does not qualify."""
        import os

        from .. import distributed as dist
        from ..ops import functional as OF

        if os.environ.get("PDT_SGD_RIDER", "1") == "0":
            return False
        group_ = getattr(model, "process_group", None)
        world = group_.size() if group_ is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        if world > 1:
            return False   # the gradients still have to be averaged: the reduce kernels carry the update (fuse_with_ddp)
        inner = getattr(model, "module", model)
        try:
            c1, b1, c2, b2, fc = inner.layer1[0], inner.layer1[1], inner.layer2[0], inner.layer2[1], inner.fc
            params = [c1.weight, c1.bias, b1.weight, b1.bias, c2.weight, c2.bias, fc.weight, fc.bias, b2.weight, b2.bias]
        except (AttributeError, IndexError, TypeError):
            return False
        if len(self.param_groups) != 1 or any(q is None for q in params):
            return False
        group = self.param_groups[0]
        mine = group["params"]
        if len(mine) != len(params) or {id(q) for q in mine} != {id(q) for q in params}:
            return False
        if not all(q.is_cuda and q.dtype == torch.float32 and q.is_contiguous() for q in params):
            return False
        self._rider_params = params
        self._rode = False

        def args(prev_grads):
            if getattr(self, "_fused_active", False):   # the reduce kernels carry the update (N >= 2)
                return None
            g = self.param_groups[0]
            bufs = []
            first = False
            if g["momentum"] != 0:
                states = [self.state[q] for q in params]
                have = [st.get("momentum_buffer") is not None for st in states]
                if any(have) and not all(have):
                    return None   # mixed first-step state: leave this iteration to step()
                first = not any(have)
                for q, st in zip(params, states):
                    if st.get("momentum_buffer") is None:
                        st["momentum_buffer"] = torch.zeros_like(q, memory_format=torch.contiguous_format)
                    bufs.append(st["momentum_buffer"])
            self._rode = True
            return (params, list(prev_grads), bufs, float(g["lr"]), self._lr_tensor(0, g, params[0].device), float(g["momentum"]),
                    float(g["dampening"]), float(g["weight_decay"]), bool(g["nesterov"]), bool(g["maximize"]), first)

        OF._sgd_rider = {"params": params, "args": args, "owner": self}
        return True

    def stop_riding(self) -> None:
        """Undo :meth:`ride_on_backward`."""
        from ..ops import functional as OF

        if OF._sgd_rider is not None and OF._sgd_rider.get("owner") is self:
            OF._sgd_rider = None
        self._rode = False

    # ---- DDP fusion: the update rides on the gradient reduction ----------------------------------------
    def fuse_with_ddp(self, ddp) -> "SGD":
        """Fuse this optimizer into DDP's gradient reduction.

        The reference's step is ``backward`` (→ NCCL allreduce of the bucket, launched after the last gradient)
        followed by a foreach SGD kernel (ddp_example.py:89-92).  Fused, every *reduce chunk* the reducer launches
        from the autograd hook — a contiguous piece of the bucket in grad-ready order — is ONE kernel on the comm
        stream that pushes the chunk into the peers' staging slots, crosses one device-side barrier, folds the
        ``world`` slots in rank order and applies the SGD update to the parameters the chunk belongs to
        (``Comm.allreduce_sgd``), so the reduction *and* the update of the late layers overlap the backward pass
        of the early ones.  The last chunk also carries DDP's per-step BatchNorm-buffer broadcast.
        ``step()`` then only has bookkeeping left; ``.grad`` reads as the averaged gradient, as in the reference.

        Contract: between ``backward()`` and ``step()`` the parameters are already updated — gradient clipping or a
        skipped ``step()`` are not expressible in this mode.  Takes effect on the first ``step()`` after the reducer
        has settled its bucket layout; until then (and whenever the preconditions fail) the ordinary path runs."""
        self._ddp = ddp
        return self

    def _maybe_activate_fusion(self) -> None:
        ddp = self._ddp
        if self._fused_active or ddp is None or len(self.param_groups) != 1:
            return
        group = self.param_groups[0]
        if group["maximize"] or ddp.process_group.size() == 1:
            return
        mine = [p for p in group["params"]]
        if len(mine) != len(ddp._params) or {id(p) for p in mine} != {id(p) for p in ddp._params}:
            return
        if any(p.dtype != torch.float32 for p in mine):
            return
        if group["momentum"] != 0:
            missing = [self.state[p].get("momentum_buffer") is None for p in ddp._params]
            if any(missing) and not all(missing) and group["dampening"] != 0:
                return  # "first step" is a per-parameter rule; wait until every parameter has its buffer
        if not ddp.enable_optimizer_fusion():
            return
        if group["momentum"] != 0:
            flat = torch.zeros_like(ddp.param_arena)
            for p, off in zip(ddp._params, ddp._param_offsets):
                view = flat[off:off + p.numel()].view(p.shape)
                st = self.state[p]
                if st.get("momentum_buffer") is not None:
                    view.copy_(st["momentum_buffer"])
                    st["momentum_buffer"] = view
            self._flat_momentum = flat
        self._fused_active = True
        ddp._rearm_fused_optimizer = self._arm_fused
        self._arm_fused()

    def _arm_fused(self) -> None:
        """(Re-)install the update the reducer applies with every reduce chunk of the *next* backward."""
        ddp, group = self._ddp, self.param_groups[0]
        first = False
        if group["momentum"] != 0:
            # zero buffer + "not first" is exact when dampening == 0 (b = μ·0 + g); a uniform first step sets the flag
            first = all(self.state[p].get("momentum_buffer") is None for p in ddp._params)
        self._armed = (float(group["lr"]), bool(first))
        ddp.reducer.set_fused_sgd(ddp.param_arena, self._flat_momentum, lr=float(group["lr"]),
                                  lr_tensor=self._lr_tensor(0, group, ddp.param_arena.device), momentum=float(group["momentum"]),
                                  dampening=float(group["dampening"]), weight_decay=float(group["weight_decay"]),
                                  nesterov=bool(group["nesterov"]), first_step=bool(first), bcast=ddp.tail_broadcast_buffer(),
                                  bcast_root=0)

    def _fused_step(self) -> bool:
        ddp = self._ddp
        if not (self._fused_active and ddp.reducer.fused_sgd):
            return False
        if not ddp.require_backward_grad_sync:
            return True   # no_sync(): gradients accumulate locally; the next synchronised backward reduces and applies them
        group = self.param_groups[0]
        # the reduction of this iteration's backward has already applied the update (see fuse_with_ddp)
        if group["momentum"] != 0:
            for p, off in zip(ddp._params, ddp._param_offsets):
                if self.state[p].get("momentum_buffer") is None:
                    self.state[p]["momentum_buffer"] = self._flat_momentum[off:off + p.numel()].view(p.shape)
        if self._armed != (float(group["lr"]), False):
            self._arm_fused()   # first-step flag drops after one update; a changed learning rate is picked up
        return True

    def load_state_dict(self, state_dict) -> None:
        """Standard behaviour, plus: when the step is fused with DDP the momentum lives in ONE flat buffer that
        mirrors the bucket — restored values are copied into it instead of replacing its views."""
        super().load_state_dict(state_dict)
        if self._flat_momentum is not None and self._ddp is not None:
            for p, off in zip(self._ddp._params, self._ddp._param_offsets):
                st = self.state.get(p)
                if st is None or st.get("momentum_buffer") is None:
                    continue
                view = self._flat_momentum[off:off + p.numel()].view(p.shape)
                if st["momentum_buffer"].data_ptr() != view.data_ptr():
                    view.copy_(st["momentum_buffer"])
                    st["momentum_buffer"] = view

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

        if getattr(self, "_rode", False):
            self._rode = False   # the last backward kernel applied this step's update (ride_on_backward)
            return loss
        if self._ddp is not None:
            if self._fused_step():
                return loss
            self._maybe_activate_fusion()  # takes effect from the next backward on
        for gi, group in enumerate(self.param_groups):
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
                if momentum == 0:
                    ops.sgd_step(params, grads, None, lr=group["lr"], momentum=0.0, dampening=group["dampening"],
                                 weight_decay=group["weight_decay"], nesterov=group["nesterov"], maximize=group["maximize"],
                                 first_step=False, lr_tensor=self._lr_tensor(gi, group, params[0].device))
                    continue
                # "first step" (buf = g, no dampening) is a per-parameter decision, as in torch: parameters that see
                # their first gradient now go through one launch with first_step=True, the rest through another
                fresh = [st["momentum_buffer"] is None for st in bufs]   # decided before any buffer is created below
                for want_first in (True, False):
                    sel = [i for i, f in enumerate(fresh) if f == want_first]
                    if not sel:
                        continue
                    ps, gs_ = [params[i] for i in sel], [grads[i] for i in sel]
                    if want_first:
                        for i in sel:
                            bufs[i]["momentum_buffer"] = torch.zeros_like(params[i], memory_format=torch.contiguous_format)
                    ops.sgd_step(ps, gs_, [bufs[i]["momentum_buffer"] for i in sel], lr=group["lr"], momentum=momentum,
                                 dampening=group["dampening"], weight_decay=group["weight_decay"], nesterov=group["nesterov"],
                                 maximize=group["maximize"], first_step=want_first,
                                 lr_tensor=self._lr_tensor(gi, group, params[0].device))
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
