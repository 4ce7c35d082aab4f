"""Phase timeline of the cooperative fused ConvNet kernels (globaltimer stamps written by thread 0 of every CTA).
    python tools/fused_trace.py            # one eager training step at batch 100, prints per-phase medians in µs
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pytorch_distributed_train_b200 as pdt
from pytorch_distributed_train_b200 import _C

B = int(os.environ.get("B", "100"))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = pdt.models.ConvNet().to(dev)
crit = pdt.nn.CrossEntropyLoss()
x = torch.rand(B, 1, 28, 28, device=dev)
y = torch.randint(0, 10, (B,), device=dev)
for _ in range(3):
    for p in net.parameters():
        p.grad = None
    crit(net(x), y).backward()
torch.cuda.synchronize()
_C.fused_convnet_trace_enable(True)
for p in net.parameters():
    p.grad = None
crit(net(x), y).backward()
t = _C.fused_convnet_trace_read()[:, :B, :].double()
_C.fused_convnet_trace_enable(False)
names = {0: ("l1_fwd", ["start", "conv done", "partial written", "barrier passed", "folded", "end"]),
         1: ("l1_bwd (+conv2 wgrad)", ["start", "partial written", "barrier passed", "folded", "conv1 wgrad partial written", "barrier 2 passed", "end",
                                       "conv2 wgrad read out of TMEM"]),
         2: ("l2_fwd", ["start", "B built + sync", "epilogue done", "partial written", "barrier passed", "folded", "pooled out written", "end"]),
         3: ("l2_bwd", ["start", "B built", "partial written", "barrier passed", "folded", "dy written", "end"])}
for k, (name, phases) in names.items():
    tk = t[k]
    if tk[:, 0].max() == 0:
        continue   # kernel did not run in this configuration
    t0 = tk[:, 0].min()
    print(f"== {name}: kernel span {(tk.max() - t0) / 1e3:.2f} us (first CTA start -> last CTA end)")
    order = sorted(range(len(phases)), key=lambda i: tk[:, i].median().item())
    for i in order:
        if tk[:, i].max() == 0:
            continue
        col = (tk[:, i] - t0) / 1e3
        print(f"   {phases[i]:30s} min {col.min():7.2f}  median {col.median():7.2f}  max {col.max():7.2f}")
