"""A few eager training steps of the ConvNet (batch 100, fused layers) in one process — the target of ncu / compute-sanitizer.
    python tools/one_step.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pytorch_distributed_train_b200 as pdt

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = pdt.models.ConvNet().to(dev)
crit = pdt.nn.CrossEntropyLoss()
opt = pdt.optim.SGD(net.parameters(), 1e-2)
x = torch.rand(100, 1, 28, 28, device=dev)
y = torch.randint(0, 10, (100,), device=dev)
from pytorch_distributed_train_b200.ops import functional as OF

losses = []
for _ in range(steps):
    # the engine's step (engine/graphed_step.py::_eager_step): targets announced before the forward pass, so the forward kernel
    # also produces the loss and its gradient — the kernels launched here are the ones a captured step replays
    with OF.upcoming_targets(y, loss_read_after_backward=True):
        out = net(x)
    loss = crit(out, y)
    opt.zero_grad()
    loss.backward()
    opt.step()
    losses.append(loss.item())
torch.cuda.synchronize()
print("losses", losses, "kernels launched", pdt._C.kernel_launch_count())
