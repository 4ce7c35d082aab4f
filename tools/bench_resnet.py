#!/usr/bin/env python
"""BASELINE.json config 4: ResNet-18 on synthetic 3×224×224, DDP (larger buckets, overlap stress).

    python -m torch.distributed.run --nproc-per-node N tools/bench_resnet.py --impl ours|reference [--syncbn]

Both arms use the same model definition (models.resnet18) and cuDNN/cuBLAS compute; what differs is the
DDP machinery: ours = C++ reducer + NVLink two-shot/NVLS bucket allreduce on the symmetric heap (+ our
SyncBatchNorm kernels), reference = torch DistributedDataParallel + NCCL (+ torch SyncBatchNorm).
Device-timed (CUDA events), max over ranks, one JSON line from rank 0.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--syncbn", action="store_true")
    ap.add_argument("--bucket-mb", type=float, default=25.0)
    a = ap.parse_args()
    import torch
    import torch.nn as nn

    import pytorch_distributed_train_b200 as pdt

    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    model = pdt.models.resnet18(num_classes=1000)
    if a.impl == "ours":
        init = "env://" if "MASTER_ADDR" in os.environ else "tcp://127.0.0.1:29631"
        pdt.init_process_group("nccl", init_method=init, world_size=world, rank=rank)
        if a.syncbn:
            model = pdt.SyncBatchNorm.convert_sync_batchnorm(model)
        model.to(dev)
        opt = pdt.optim.SGD(model.parameters(), 0.01, momentum=0.9)
        ddp = pdt.DistributedDataParallel(model, device_ids=[local], bucket_cap_mb=a.bucket_mb)
        barrier = pdt.distributed.barrier
        from pytorch_distributed_train_b200.utils import max_over_ranks
    else:
        import torch.distributed as dist

        port = int(os.environ.get("MASTER_PORT", "29500")) + 21
        os.environ.pop("TORCHELASTIC_USE_AGENT_STORE", None)  # else every rank would be a client of our tcp:// port
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
        if a.syncbn:
            model = nn.SyncBatchNorm.convert_sync_batchnorm(model)
        model.to(dev)
        opt = torch.optim.SGD(model.parameters(), 0.01, momentum=0.9)
        ddp = nn.parallel.DistributedDataParallel(model, device_ids=[local], bucket_cap_mb=a.bucket_mb)
        barrier = dist.barrier

        def max_over_ranks(v):
            t = torch.tensor([v], device=dev)
            dist.all_reduce(t, dist.ReduceOp.MAX)
            return float(t.item())
    crit = nn.CrossEntropyLoss()
    x = torch.randn(a.batch, 3, 224, 224, device=dev)
    y = torch.randint(0, 1000, (a.batch,), device=dev)

    def step():
        loss = crit(ddp(x), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1))
    if rank == 0:
        out = {"model": "resnet18", "impl": a.impl, "n_gpus": world, "per_gpu_batch": a.batch, "syncbn": a.syncbn,
               "ms_per_step": ms / a.steps, "images_per_s": a.batch * world * a.steps / (ms / 1e3), "loss": float(loss.detach())}
        if a.impl == "ours":
            info = ddp._get_ddp_logging_data()
            out.update({"buckets": info["bucket_sizes"], "comm": info["comm_kind"], "copies_into_bucket": info["copies_into_bucket"],
                        "bwd_comm_exposed_us": info["backward_comm_exposed_us"]})
        print(json.dumps(out), flush=True)
    if a.impl == "ours":
        pdt.destroy_process_group()
    else:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
