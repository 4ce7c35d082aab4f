"""Single-process multi-GPU ``DataParallel`` — the strategy the reference's README describes
(but never runs) next to DDP (ref: README.md:10-17; torch/nn/parallel/data_parallel.py:54,173-220):
scatter the *global* batch along ``dim`` across ``device_ids``, run a replica per device on its
own thread, gather outputs on ``output_device``.  Gradients flow back to the original parameters
because replicas are built with differentiable broadcasts of the source parameters.

Kept deliberately small: DistributedDataParallel is the product; this exists for API parity.
"""
from __future__ import annotations

import threading
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
from torch.func import functional_call


def _scatter(obj, devices, dim):
    if isinstance(obj, torch.Tensor):
        chunks = obj.chunk(len(devices), dim)
        return [c.to(d, non_blocking=True) for c, d in zip(chunks, devices)]
    if isinstance(obj, (tuple, list)):
        per = [_scatter(o, devices, dim) for o in obj]
        n = min(len(p) for p in per) if per else len(devices)
        return [type(obj)(p[i] for p in per) for i in range(n)]
    if isinstance(obj, dict):
        per = {k: _scatter(v, devices, dim) for k, v in obj.items()}
        n = min(len(p) for p in per.values()) if per else len(devices)
        return [{k: p[i] for k, p in per.items()} for i in range(n)]
    return [obj for _ in devices]


def _gather(outs, device, dim):
    o0 = outs[0]
    if isinstance(o0, torch.Tensor):
        return torch.cat([o.to(device) for o in outs], dim)
    if isinstance(o0, (tuple, list)):
        return type(o0)(_gather([o[i] for o in outs], device, dim) for i in range(len(o0)))
    if isinstance(o0, dict):
        return {k: _gather([o[k] for o in outs], device, dim) for k in o0}
    return o0


class DataParallel(nn.Module):
    def __init__(self, module: nn.Module, device_ids: Optional[Sequence] = None, output_device=None, dim: int = 0):
        super().__init__()
        self.module = module
        self.dim = dim
        if device_ids is None:
            device_ids = list(range(torch.cuda.device_count())) if torch.cuda.is_available() else []
        self.device_ids = [torch.device("cuda", d) if isinstance(d, int) else torch.device(d) for d in device_ids]
        self.output_device = (torch.device("cuda", output_device) if isinstance(output_device, int)
                              else torch.device(output_device) if output_device is not None
                              else (self.device_ids[0] if self.device_ids else None))

    def forward(self, *inputs, **kwargs):
        if len(self.device_ids) <= 1:
            return self.module(*inputs, **kwargs)
        ins = _scatter(tuple(inputs), self.device_ids, self.dim)
        kws = _scatter(kwargs, self.device_ids, self.dim) if kwargs else [{} for _ in ins]
        devices = self.device_ids[:len(ins)]
        names = [n for n, _ in self.module.named_parameters()] + [n for n, _ in self.module.named_buffers()]
        tensors = list(self.module.parameters()) + list(self.module.buffers())
        outs: List = [None] * len(ins)
        errs: List = [None] * len(ins)

        def run(i):
            try:
                dev = devices[i]
                with torch.cuda.device(dev) if dev.type == "cuda" else torch.device("cpu"):
                    replica = {n: t.to(dev) for n, t in zip(names, tensors)}  # differentiable copy
                    outs[i] = functional_call(self.module, replica, ins[i], kws[i])
            except BaseException as e:  # noqa: BLE001
                errs[i] = e

        threads = [threading.Thread(target=run, args=(i,)) for i in range(1, len(ins))]
        for t in threads:
            t.start()
        run(0)
        for t in threads:
            t.join()
        for e in errs:
            if e is not None:
                raise e
        return _gather(outs, self.output_device, self.dim)
