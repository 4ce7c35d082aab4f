"""Training entry point — the framework's equivalent of the reference script
(ref: ddp_example.py:47-115).  Same four flags with the same names and defaults
(``-g/--gpus``, ``--epochs``, ``--backend``, ``--syncbn``; ref: ddp_example.py:103-106), same
stdout lines (``Rank id:``, ``Use SyncBN in training``, ``Epoch [e/E], Step [i/N], Loss: x``,
``Training complete in:``; ref: ddp_example.py:49,56,94,97), same flow: spawn one process per
GPU → init_process_group over TCP → seed → model → optional SyncBN → DDP → sampler/loader → loop.

Extra flags cover what the reference hard-codes: ``--init-method`` (its LAN address
``tcp://10.9.1.2:34567`` only works on the author's network, ref: ddp_example.py:110; we default
to loopback with a free port), ``--data synthetic|mnist``, ``--model``, ``--comm fused|nccl``, ``--algo``,
``--steps``, ``--graph`` (whole-step CUDA graph), ``--batch-size``, ``--lr``, ``--checkpoint`` / ``--resume``.
"""
from __future__ import annotations

import argparse
import socket
import sys
from datetime import datetime

import torch


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="B200-native DDP training (MNIST ConvNet / ResNet-18)")
    p.add_argument("-g", "--gpus", default=1, type=int, help="number of gpus per node")
    p.add_argument("--epochs", default=2, type=int, metavar="N", help="number of total epochs to run")
    p.add_argument("--backend", default="nccl", type=str, help="backend used for distributed train (nccl | gloo)")
    p.add_argument("--syncbn", default=False, action="store_true", help="whether to use syncbn while training")
    p.add_argument("--init-method", default=None, type=str, help="rendezvous URL (default: tcp://127.0.0.1:<free port>)")
    p.add_argument("--comm", default="fused", choices=["fused", "nccl"],
                   help="GPU collectives: our fused NVLink kernels (default) or the libnccl baseline")
    p.add_argument("--algo", default="auto", choices=["auto", "oneshot", "oneshot_mc", "twoshot", "nvls"],
                   help="allreduce algorithm of the fused backend (auto: by message size and world size; sets PDT_AR_ALGO)")
    p.add_argument("--data", default="synthetic", choices=["synthetic", "mnist"], help="dataset (no network: synthetic default)")
    p.add_argument("--data-root", default="./data", type=str)
    p.add_argument("--model", default="convnet", choices=["convnet", "resnet18"])
    p.add_argument("--batch-size", default=100, type=int, help="per-GPU batch size (reference: 100)")
    p.add_argument("--lr", default=1e-4, type=float, help="SGD learning rate (reference: 1e-4)")
    p.add_argument("--momentum", default=0.0, type=float)
    p.add_argument("--steps", default=0, type=int, help="stop each epoch after this many steps (0 = full epoch)")
    p.add_argument("--samples", default=60000, type=int, help="synthetic dataset size")
    p.add_argument("--graph", default=False, action="store_true", help="capture the whole training step in a CUDA graph")
    p.add_argument("--log-interval", default=10, type=int)
    p.add_argument("--checkpoint", default=None, type=str, help="write a checkpoint here at the end of every epoch (rank 0, atomic)")
    p.add_argument("--resume", default=None, type=str, help="restore model/optimizer/epoch from this checkpoint before training")
    p.add_argument("--set-epoch", default=False, action="store_true",
                   help="call sampler.set_epoch(e) each epoch (the reference does not)")
    return p


def dist_train(gpu: int, args) -> None:
    """Per-process body: ``fn(i, *args)`` target of the launcher."""
    import pytorch_distributed_train_b200 as pdt
    from pytorch_distributed_train_b200 import data as pdata

    rank = gpu  # single node: spawn index is both global rank and device ordinal (ref: ddp_example.py:48)
    print("Rank id: ", rank)
    use_cuda = args.backend not in ("gloo", "cpu")
    if use_cuda:
        torch.cuda.set_device(gpu)
    if getattr(args, "algo", "auto") != "auto":
        import os

        os.environ["PDT_AR_ALGO"] = args.algo   # read by the fused backend when the process group is created
    pdt.init_process_group(backend=args.backend, init_method=args.init_method, world_size=args.world_size,
                           rank=rank, comm=args.comm)
    torch.manual_seed(0)
    if args.model == "convnet":
        model = pdt.models.ConvNet()
        shape = (1, 28, 28)
    else:
        model = pdt.models.resnet18(num_classes=1000)
        shape = (3, 224, 224)
    if args.syncbn:
        model = pdt.SyncBatchNorm.convert_sync_batchnorm(model)
        if gpu == 0:
            print("Use SyncBN in training")
    device = torch.device("cuda", gpu) if use_cuda else torch.device("cpu")
    model.to(device)
    batch_size = args.batch_size
    criterion = pdt.nn.CrossEntropyLoss().to(device)
    optimizer = pdt.optim.SGD(model.parameters(), args.lr, momentum=args.momentum)
    model = pdt.DistributedDataParallel(model, device_ids=[gpu] if use_cuda else None)

    if args.data == "mnist" and args.model == "convnet":
        train_dataset = pdata.MNIST(root=args.data_root, train=True, download=True, synthetic_fallback=True)
    else:
        n = args.samples if args.model == "convnet" else min(args.samples, 4096)
        train_dataset = pdata.SyntheticMNIST(n, seed=0, num_classes=10 if args.model == "convnet" else 1000, image_shape=shape)
    train_sampler = pdt.DistributedSampler(train_dataset, num_replicas=args.world_size, rank=rank)
    train_loader = pdt.DataLoader(dataset=train_dataset, batch_size=batch_size, shuffle=False, num_workers=0,
                                  pin_memory=use_cuda, sampler=train_sampler)

    step_fn = None
    if args.graph and use_cuda:
        from pytorch_distributed_train_b200.engine import GraphedTrainStep

        step_fn = GraphedTrainStep(model, criterion, optimizer, example_inputs=(
            torch.zeros((batch_size,) + shape, device=device), torch.zeros(batch_size, dtype=torch.int64, device=device)))

    first_epoch = 0
    if args.resume:
        info = pdt.utils.load_checkpoint(args.resume, model, optimizer, sampler=train_sampler)
        first_epoch = info["epoch"]
        if gpu == 0:
            print(f"Resumed from {args.resume} at epoch {first_epoch}")

    start = datetime.now()
    total_step = len(train_loader)
    for epoch in range(first_epoch, args.epochs):
        if args.set_epoch:
            train_sampler.set_epoch(epoch)
        pending = None   # (step index, loss handle) of a log line whose value is still on its way to the host
        fmt = "Epoch [{}/{}], Step [{}/{}], Loss: {:.4f}"
        for i, (images, labels) in enumerate(train_loader):
            if args.steps and i >= args.steps:
                break
            graphed = step_fn is not None and images.shape[0] == batch_size
            if graphed:
                # the (pinned) host batch goes straight into the captured step's input buffers; the copy overlaps the previous step
                step_fn(images, labels)
                handle = step_fn.loss_to_host()
            else:
                images = images.to(device, non_blocking=True)
                labels = labels.to(device, non_blocking=True)
                outputs = model(images)
                loss = criterion(outputs, labels)
                optimizer.zero_grad()
                loss.backward()
                optimizer.step()
                handle = loss
            if pending is not None and gpu == 0:
                # the log line of the previous step, printed once this step has been queued: the GPU keeps working while the host waits
                print(fmt.format(epoch + 1, args.epochs, pending[0], total_step, pending[1].item()))
                pending = None
            if (i + 1) % args.log_interval == 0 and gpu == 0:
                if graphed:
                    pending = (i + 1, handle)
                else:
                    print(fmt.format(epoch + 1, args.epochs, i + 1, total_step, handle.item()))
        if pending is not None and gpu == 0:
            print(fmt.format(epoch + 1, args.epochs, pending[0], total_step, pending[1].item()))
        if args.checkpoint:
            pdt.utils.save_checkpoint(args.checkpoint, model, optimizer, epoch=epoch + 1, sampler=train_sampler)
    if use_cuda:
        torch.cuda.synchronize()
    if gpu == 0:
        print("Training complete in: " + str(datetime.now() - start))
    pdt.destroy_process_group()


def main(argv=None) -> None:
    args = build_parser().parse_args(argv)
    args.world_size = args.gpus  # one process per GPU (ref: ddp_example.py:109)
    if args.init_method is None:
        args.init_method = f"tcp://127.0.0.1:{_free_port()}"
    from pytorch_distributed_train_b200.launcher import spawn

    spawn(dist_train, nprocs=args.gpus, args=(args,))


if __name__ == "__main__":
    main(sys.argv[1:])
