#!/usr/bin/env python
"""BASELINE.json config 1: MNIST ConvNet DDP, world_size=2, CPU (no GPU needed) — our C++ store + TCP-mesh CPU backend +
C++ reducer versus torch DistributedDataParallel + gloo, same model code (torch CPU kernels on both sides), same batches.

    python tools/bench_cpu_gloo.py [--steps 30] [--world 2]

Wall-clock per step (max over ranks), after warm-up; prints one JSON line per arm.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def worker(rank, world, steps, warmup, port):
    import torch
    import torch.distributed as td
    import torch.nn as nn

    import pytorch_distributed_train_b200 as pdt

    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))

    def data(s):
        g = torch.Generator().manual_seed(1000 * s + rank)
        return torch.rand(100, 1, 28, 28, generator=g), torch.randint(0, 10, (100,), generator=g)

    batches = [data(s) for s in range(8)]

    def run(ddp, opt, crit, barrier):
        for s in range(warmup):
            x, y = batches[s % 8]
            loss = crit(ddp(x), y); opt.zero_grad(); loss.backward(); opt.step()
        barrier()
        t0 = time.perf_counter()
        for s in range(steps):
            x, y = batches[s % 8]
            loss = crit(ddp(x), y); opt.zero_grad(); loss.backward(); opt.step()
        barrier()
        return (time.perf_counter() - t0) / steps * 1e3, loss.item()

    torch.manual_seed(0)
    m = pdt.models.ConvNet()
    ours_ms, ours_loss = run(pdt.DistributedDataParallel(m), pdt.optim.SGD(m.parameters(), 1e-4), pdt.nn.CrossEntropyLoss(),
                             pdt.distributed.barrier)
    td.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
    torch.manual_seed(0)
    m = pdt.models.ConvNet(fused=False)
    ref_ms, ref_loss = run(nn.parallel.DistributedDataParallel(m), torch.optim.SGD(m.parameters(), 1e-4), nn.CrossEntropyLoss(), td.barrier)
    t = torch.tensor([ours_ms, ref_ms])
    td.all_reduce(t, td.ReduceOp.MAX)
    td.destroy_process_group()
    return t.tolist(), ours_loss, ref_loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--world", type=int, default=2)
    a = ap.parse_args()
    from mp_helpers import free_port, run_ranks

    (ours_ms, ref_ms), lo, lr = run_ranks(worker, a.world, a.steps, a.warmup, free_port())[0]
    for impl, ms, loss in (("ours (C++ store + CPU backend + reducer)", ours_ms, lo), ("torch DDP + gloo", ref_ms, lr)):
        print(json.dumps({"config": f"ConvNet DDP world_size={a.world} CPU", "impl": impl, "ms_per_step": round(ms, 2),
                          "images_per_s": round(100 * a.world / (ms / 1e3)), "final_loss": round(loss, 5)}))


if __name__ == "__main__":
    main()
