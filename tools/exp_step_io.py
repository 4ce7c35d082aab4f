"""Where does the host time of an end-to-end step go?  One rank, the bench's e2e loop taken apart: per-call host times of the loader, the
native input staging, the graph replay and the loss hand-off, with loader batches and with one fixed pinned batch."""
import os
import socket
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pytorch_distributed_train_b200 as pdt
from pytorch_distributed_train_b200.engine import GraphedTrainStep

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
with socket.socket() as s_:
    s_.bind(("127.0.0.1", 0))
    port = s_.getsockname()[1]
pdt.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", world_size=1, rank=0)
torch.manual_seed(0)
net = pdt.models.ConvNet().to(dev)
opt = pdt.optim.SGD(net.parameters(), 1e-4)
ddp = pdt.DistributedDataParallel(net, device_ids=[0])
crit = pdt.nn.CrossEntropyLoss()
x = torch.rand(100, 1, 28, 28, device=dev)
y = torch.randint(0, 10, (100,), device=dev)
step = GraphedTrainStep(ddp, crit, opt, (x, y))
ds = pdt.data.MNIST(root="/tmp/exp_mnist", train=True, synthetic_fallback=True)   # uint8 idx files, like bench.py's e2e arm
loader = pdt.DataLoader(ds, batch_size=100, pin_memory=True, sampler=pdt.DistributedSampler(ds, 1, 0), prefetch=8)


def batches():
    while True:
        for b in loader:
            yield b


def run(name, get, n=400, read_every=10):
    it = get()
    T = {"next": [], "stage+replay": [], "loss_to_host": [], "read": []}
    pending = None
    for _ in range(30):
        xb, yb = next(it)
        step(xb, yb)
        step.loss_to_host()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True)
    b = torch.cuda.Event(enable_timing=True)
    a.record()
    w0 = time.perf_counter()
    for i in range(n):
        t0 = time.perf_counter()
        xb, yb = next(it)
        t1 = time.perf_counter()
        gi = step.replays % len(step.graphs)
        step._io.stage_inputs(gi, step.input_sets[gi % len(step.input_sets)], [xb, yb], False, step.double_buffer)
        tm = time.perf_counter()
        step.graphs[gi].replay()
        step._io.replayed(gi)
        step.replays += 1
        step._last = gi
        step.static_loss = step.losses[gi]
        t2 = time.perf_counter()
        T.setdefault("stage only", []).append((tm - t1) * 1e6)
        h = step.loss_to_host()
        t3 = time.perf_counter()
        if pending is not None:
            pending.item()
            pending = None
        if read_every and (i + 1) % read_every == 0:
            pending = h
        t4 = time.perf_counter()
        for k, d in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            T[k].append(d * 1e6)
    b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - w0) * 1e6 / n
    med = {k: round(sorted(v)[len(v) // 2], 1) for k, v in T.items()}
    p90 = {k: round(sorted(v)[int(len(v) * 0.9)], 1) for k, v in T.items()}
    if getattr(loader, "_native", None) is not None and "loader" in name:
        st = loader._native.stats()
        print(f"   stager: {st[0]:.0f} batches, fill {st[1] / max(st[0], 1):.1f} us/batch, waited for a free slot {st[2] / max(st[0], 1):.1f} us/batch; "
              f"next(): {st[3]:.0f} calls, {st[4] / max(st[3], 1):.2f} ready slots on entry, waited {st[5] / max(st[3], 1):.1f} us/call; "
              f"slots re-allocated {st[6]:.0f} ({st[7] / max(st[0], 1):.1f} us/batch)", flush=True)
    print(f"{name:34s} device {a.elapsed_time(b) * 1e3 / n:7.1f} us/step  wall {wall:7.1f}  host median {med}  p90 {p90}", flush=True)


xp, yp = x.cpu().pin_memory(), y.cpu().pin_memory()


def fixed():
    while True:
        yield xp, yp


def fixed_dev():
    while True:
        yield x, y


ring = [(torch.rand(100, 1, 28, 28).pin_memory(), torch.randint(0, 10, (100,)).pin_memory()) for _ in range(12)]


def fixed_ring():
    k = 0
    while True:
        yield ring[k % 12]
        k += 1


xb0, yb0 = next(iter(loader))
print("loader batch pinned:", xb0.is_pinned(), yb0.is_pinned(), xb0.dtype, yb0.dtype, xb0.data_ptr() % 4096, flush=True)
run("device-resident batch", fixed_dev)
run("ring of 12 fixed pinned batches", fixed_ring)
run("one fixed pinned batch", fixed)
run("loader (native stager, pinned)", batches)
run("loader, no loss reads", batches, read_every=0)
# 20-batch epochs: an epoch boundary every 20 steps (8 ranks on MNIST: every 75)
ds_small = pdt.data.SyntheticMNIST(2000)
loader = pdt.DataLoader(ds_small, batch_size=100, pin_memory=True, sampler=pdt.DistributedSampler(ds_small, 1, 0), prefetch=8)
run("loader, 20-batch epochs", batches)
os.environ["X"] = "1"
pdt.destroy_process_group()
