"""Who keeps a loader batch alive?  Counts slot re-allocations of the native stager under different consumers."""
import gc
import os
import sys
import weakref

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pytorch_distributed_train_b200 as pdt
from pytorch_distributed_train_b200 import _C

dev = torch.device("cuda", 0)
ds = pdt.data.MNIST(root="/tmp/exp_mnist", train=True, synthetic_fallback=True)
xd = torch.empty(100, 1, 28, 28, device=dev)
yd = torch.empty(100, dtype=torch.int64, device=dev)
io = _C.StepPipeline(0, 2, torch.zeros((), device=dev))


def trial(name, consume, n=200):
    loader = pdt.DataLoader(ds, batch_size=100, pin_memory=True, sampler=pdt.DistributedSampler(ds, 1, 0), prefetch=8)
    refs = []
    for k, (xb, yb) in enumerate(loader):
        refs.append(weakref.ref(xb))
        consume(k, xb, yb)
        if k + 1 == n:
            break
    alive = sum(r() is not None for r in refs[:-2])
    st = loader._native.stats()
    print(f"{name:44s} re-allocated {st[6]:.0f}/{st[0]:.0f} slots; python tensors of old batches still alive: {alive}", flush=True)
    torch.cuda.synchronize()


trial("iterate only", lambda k, xb, yb: None)
trial("python copy_ (current stream)", lambda k, xb, yb: (xd.copy_(xb, non_blocking=True), yd.copy_(yb, non_blocking=True)))
trial("StepPipeline.stage_inputs (overlap)", lambda k, xb, yb: (io.stage_inputs(k % 2, [xd, yd], [xb, yb], False, True), io.replayed(k % 2)))
trial("StepPipeline.stage_inputs (same stream)", lambda k, xb, yb: io.stage_inputs(k % 2, [xd, yd], [xb, yb], False, False))
