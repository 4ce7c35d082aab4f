#!/usr/bin/env python
"""Headline benchmark: MNIST ConvNet DDP training throughput (images/sec, device-timed, max over
ranks) — the metric and config BASELINE.json names (ref: ddp_example.py ConvNet, batch 100/GPU,
SGD lr 1e-4, fp32 with TF32 tensor-core convolutions, synthetic MNIST-shaped data, random init).

    python bench.py --gpus N --steps K --warmup W                 # this framework
    python bench.py --impl reference --gpus N --steps K --warmup W # unmodified reference stack

N > 1 is launched one rank per GPU by ``python -m torch.distributed.run --nproc-per-node N ...``
(RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment).  Rank 0 prints ONE JSON line.

Timed regions (both arms): W >= 3 untimed warm-up steps, then exactly K steps between
barrier + torch.cuda.synchronize() on both sides, CUDA events on the launching stream, MAX over
ranks.  ``value`` is the device-timed step with inputs rotating through a device-resident pool that
is larger than L2 (164 MB > 126 MB).  ``e2e`` is the same step driven through the public API from
*pinned host memory* (H2D of every batch inside the timed region) with a D2H read of every loss.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 100                    # per GPU (ref: ddp_example.py:59)
IMG = (1, 28, 28)
POOL_BATCHES = 512             # 512 x 100 x 784 x 4 B = 160.6 MB of images  > 126 MB L2
LR = 1e-4                      # ref: ddp_example.py:62


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=20)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--syncbn", action="store_true")
    p.add_argument("--comm", default="fused", choices=["fused", "nccl"])
    p.add_argument("--no-graph", action="store_true", help="eager steps instead of the whole-step CUDA graph")
    p.add_argument("--conv-impl", default="auto", choices=["auto", "simt", "tcgen05"])
    p.add_argument("--skip-e2e", action="store_true")
    return p.parse_args()


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def clock_block(sampler):
    s = sampler.summary()
    return {"sm_mhz": s.get("sm_mhz"), "sm_max_mhz": s.get("sm_max_mhz"), "reasons": s.get("reasons", []),
            "power_w_max": s.get("power_w_max"), "samples": s.get("samples", 0)}


# =====================================================================================================
# this framework
# =====================================================================================================
def run_ours(args):
    import torch

    import pytorch_distributed_train_b200 as pdt
    from pytorch_distributed_train_b200 import _C
    from pytorch_distributed_train_b200.engine import GraphedTrainStep
    from pytorch_distributed_train_b200.utils import ClockSampler, max_over_ranks

    rank, local_rank, world = env_rank()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if "MASTER_ADDR" not in os.environ:
        import socket

        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        init = f"tcp://127.0.0.1:{port}"
    else:
        init = "env://"
    pdt.init_process_group(backend="nccl", init_method=init, world_size=world, rank=rank, comm=args.comm)
    group = pdt.distributed.get_default_group()

    torch.manual_seed(0)  # identical init on every rank (ref: ddp_example.py:51)
    model = pdt.models.ConvNet()
    if args.conv_impl != "auto":
        os.environ["PDT_CONV_IMPL"] = args.conv_impl
    if args.syncbn:
        model = pdt.SyncBatchNorm.convert_sync_batchnorm(model)
    model.to(dev)
    criterion = pdt.nn.CrossEntropyLoss().to(dev)
    optimizer = pdt.optim.SGD(model.parameters(), LR)
    ddp = pdt.DistributedDataParallel(model, device_ids=[local_rank])

    # synthetic MNIST-shaped data: a host pool in pinned memory and a device copy of it
    g = torch.Generator().manual_seed(1234 + rank)
    host_x = torch.rand((POOL_BATCHES, BATCH) + IMG, generator=g).pin_memory()
    host_y = torch.randint(0, 10, (POOL_BATCHES, BATCH), generator=g).pin_memory()
    dev_x, dev_y = host_x.to(dev), host_y.to(dev)

    if args.no_graph:
        def step(x, y):
            x = x.to(dev, non_blocking=True)
            y = y.to(dev, non_blocking=True)
            loss = criterion(ddp(x), y)
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            return loss
        launches_per_step = None
    else:
        before = _C.kernel_launch_count()
        graphed = GraphedTrainStep(ddp, criterion, optimizer, (dev_x[0], dev_y[0]), warmup=3, zero_grad_set_to_none=True,
                                   fuse_optimizer=os.environ.get("PDT_FUSE_OPT", "1") != "0",
                                   double_buffer_inputs=os.environ.get("PDT_E2E_DOUBLE_BUFFER", "0") == "1")
        launches_per_step = graphed.kernels_per_replay
        step = graphed
        del before

    W, K = max(args.warmup, 3), args.steps

    def timed(run_step, n_warm, n_timed):
        for i in range(n_warm):
            run_step(i)
        torch.cuda.synchronize()
        pdt.distributed.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = _C.kernel_launch_count()
        t0 = time.perf_counter()
        a.record()
        for i in range(n_timed):
            run_step(n_warm + i)
        b.record()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        pdt.distributed.barrier()
        torch.cuda.synchronize()
        return a.elapsed_time(b), wall_ms, _C.kernel_launch_count() - c0

    # ---- device-timed steps, inputs rotate through a device pool larger than L2 ------------------------
    last = {}

    def dev_step(i):
        j = i % POOL_BATCHES
        last["loss"] = step(dev_x[j], dev_y[j])

    with ClockSampler(gpu_index=local_rank, period_ms=100) as clocks:
        ms_dev, _, eager_launches = timed(dev_step, W, K)
        # ---- end to end: pinned host batches in, every loss read back ---------------------------------
        e2e = None
        if not args.skip_e2e:
            R = 4
            host_loss = torch.zeros(R, dtype=torch.float32).pin_memory()
            evs = [torch.cuda.Event() for _ in range(R)]

            def e2e_step(i):
                j = i % POOL_BATCHES
                if i >= R:
                    evs[i % R].synchronize()          # the slot's previous loss has reached the host
                loss = step(host_x[j], host_y[j])     # H2D of this step's batch happens inside the call
                host_loss[i % R].copy_(loss.detach(), non_blocking=True)
                evs[i % R].record()

            ms_e2e_dev, ms_e2e_wall, _ = timed(e2e_step, W, K)
            ms_e2e = max(ms_e2e_dev, ms_e2e_wall)
            e2e = {"ms": ms_e2e, "loss": float(host_loss[(W + K - 1) % R])}
    ms_dev = max_over_ranks(ms_dev)
    out = None
    if e2e is not None:
        e2e["ms"] = max_over_ranks(e2e["ms"])
    if rank == 0:
        imgs = BATCH * world * K
        value = imgs / (ms_dev / 1e3)
        info = ddp._get_ddp_logging_data()
        gpu_launches = (launches_per_step * K) if launches_per_step is not None else eager_launches
        out = {
            "metric": "MNIST ConvNet DDP training throughput (images/sec, device-timed, max over ranks)",
            "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32 (tf32 tensor-core conv, fp32 accumulate) — the reference's precision",
            "data": "synthetic MNIST-shaped images, random-init weights",
            "impl": "ours",
            "config": {"model": "ConvNet (ref ddp_example.py:22-41)", "global_batch": BATCH * world, "per_gpu_batch": BATCH,
                       "seq_len": None, "parallelism": f"dp{world}", "optimizer": "SGD lr=1e-4", "syncbn": args.syncbn,
                       "comm": info.get("comm_kind"), "cuda_graph": not args.no_graph,
                       "inputs": f"rotating pool of {POOL_BATCHES} batches = {dev_x.numel() * 4 / 1e6:.0f} MB > 126 MB L2 (no explicit L2 flush)",
                       "buckets": info.get("bucket_sizes"), "grad_copies_into_bucket": info.get("copies_into_bucket"),
                       "fused_allreduce_sgd": bool(getattr(optimizer, "_fused_active", False))},
            "gpu_launches": int(gpu_launches),
            "gpu_launches_per_step": launches_per_step if launches_per_step is not None else eager_launches / K,
            "clocks": clock_block(clocks),
            "final_loss": float(last["loss"].detach()),
        }
        if e2e is not None:
            out["e2e"] = {"value": imgs / (e2e["ms"] / 1e3), "unit": "images/s", "ms_per_step": e2e["ms"] / K,
                          "h2d_bytes_per_step": BATCH * 784 * 4 + BATCH * 8, "d2h_bytes_per_step": 4,
                          "note": "GraphedTrainStep(images_pinned_cpu, labels_pinned_cpu): H2D + graph replay + async D2H of the loss, every step"}
    pdt.destroy_process_group()
    return out


# =====================================================================================================
# reference arm: the unmodified reference from baseline/_ref, its own code path
# =====================================================================================================
def _ensure_reference():
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "ddp_example.py")):
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import install_ref

        install_ref.install()
    return ref_dir


def run_reference(args):
    try:
        ref_dir = _ensure_reference()
        sys.path.insert(0, ref_dir)
        import ddp_example as ref  # noqa: F401  (unmodified copy of /root/reference/ddp_example.py)
    except Exception as e:  # noqa: BLE001
        return {"impl": "reference", "unavailable": f"reference not importable: {type(e).__name__}: {e}"[:300]}
    import torch
    import torch.distributed as dist
    import torch.nn as nn

    from pytorch_distributed_train_b200.data import synthesize_mnist_files
    from pytorch_distributed_train_b200.utils import ClockSampler

    rank, local_rank, world = env_rank()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    import faulthandler

    faulthandler.dump_traceback_later(240, repeat=True, file=sys.stderr)  # say where we are if something wedges
    W, K = max(args.warmup, 3), args.steps
    base_port = int(os.environ.get("MASTER_PORT", "29500"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(local_rank)

    # The reference rendezvouses over its own tcp:// endpoint (ddp_example.py:55); under torchrun the
    # TORCHELASTIC_USE_AGENT_STORE flag would make *every* rank a client of that port (nobody serves it).
    os.environ.pop("TORCHELASTIC_USE_AGENT_STORE", None)
    # ---- (1) stack-only: the reference's model / DDP / loss / optimizer with device-resident inputs ----
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{base_port + 11}", world_size=world, rank=rank)
    torch.manual_seed(0)
    model = ref.ConvNet()
    if args.syncbn:
        model = nn.SyncBatchNorm.convert_sync_batchnorm(model)
    model.cuda(local_rank)
    criterion = nn.CrossEntropyLoss().cuda(local_rank)
    optimizer = torch.optim.SGD(model.parameters(), LR)
    model = nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])
    g = torch.Generator().manual_seed(1234 + rank)
    dev_x = torch.rand((POOL_BATCHES, BATCH) + IMG, generator=g).to(dev)
    dev_y = torch.randint(0, 10, (POOL_BATCHES, BATCH), generator=g).to(dev)

    def ref_step(i):
        j = i % POOL_BATCHES
        outputs = model(dev_x[j])
        loss = criterion(outputs, dev_y[j])
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        return loss

    with ClockSampler(gpu_index=local_rank, period_ms=100) as clocks:
        for i in range(W):
            ref_step(i)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(K):
            loss = ref_step(W + i)
        b.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        dist.all_reduce(t, dist.ReduceOp.MAX)
        ms_dev = float(t.item())
        final_loss = float(loss.detach())
        del model, optimizer
        dist.destroy_process_group()

        # ---- (2) end to end: ddp_example.dist_train itself, untouched ---------------------------------
        # Offline substitutes for what the script fetches/assumes: synthetic idx files where torchvision's
        # MNIST expects them (./data/MNIST/raw), loopback init address.  Timing is injected by wrapping the
        # DataLoader iterator (a torch class, not reference code): sync + barrier + event at step W and W+K.
        e2e_ms = None
        e2e_err = None
        if not args.skip_e2e:
            try:
                data_root = os.path.join(os.getcwd(), "data")
                marker = os.path.join(data_root, "MNIST", "raw", ".synthetic_ready")
                if local_rank == 0 and not os.path.exists(marker):
                    synthesize_mnist_files(data_root, train=True)
                    synthesize_mnist_files(data_root, train=False)
                    open(marker, "w").close()
                while not os.path.exists(marker):
                    time.sleep(0.05)
                e2e_ms = _reference_e2e(ref, args, rank, local_rank, world, W, K, base_port + 12)
            except Exception as e:  # noqa: BLE001
                e2e_err = f"{type(e).__name__}: {e}"[:300]
    if rank != 0:
        return None
    imgs = BATCH * world * K
    out = {
        "metric": "MNIST ConvNet DDP training throughput (images/sec, device-timed, max over ranks)",
        "value": imgs / (ms_dev / 1e3), "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32 (torch defaults: TF32 allowed in cuDNN conv, fp32 matmul)", "data": "synthetic MNIST-shaped images, random-init weights",
        "impl": "reference",
        "config": {"model": "ddp_example.ConvNet", "global_batch": BATCH * world, "per_gpu_batch": BATCH, "seq_len": None,
                   "parallelism": f"dp{world}", "optimizer": "torch.optim.SGD lr=1e-4", "syncbn": args.syncbn,
                   "stack": f"torch {torch.__version__} DistributedDataParallel + NCCL {'.'.join(map(str, torch.cuda.nccl.version()))} + cuDNN/cuBLAS",
                   "inputs": f"rotating pool of {POOL_BATCHES} batches = {dev_x.numel() * 4 / 1e6:.0f} MB > 126 MB L2"},
        "gpu_launches": 0, "clocks": clock_block(clocks), "final_loss": final_loss,
    }
    if e2e_ms is not None:
        out["e2e"] = {"value": imgs / (e2e_ms / 1e3), "unit": "images/s", "ms_per_step": e2e_ms / K,
                      "h2d_bytes_per_step": BATCH * 784 * 4 + BATCH * 8, "d2h_bytes_per_step": 0.4,
                      "note": "ddp_example.dist_train() unmodified: torchvision MNIST (PIL per sample) + DistributedSampler + DataLoader(pin_memory) + .cuda(non_blocking) + loss.item() every 10 steps"}
    elif e2e_err:
        out["e2e_error"] = e2e_err
    return out


def _reference_e2e(ref, args, rank, local_rank, world, W, K, port):
    import io
    from contextlib import redirect_stdout
    from types import SimpleNamespace

    import torch
    import torch.distributed as dist
    import torch.utils.data as tud

    steps_per_epoch = -(-(-(-60000 // world)) // BATCH)  # ceil(ceil(60000/world)/100): 600/300/150/75
    epochs = -(-(W + K + 1) // steps_per_epoch)
    state = {"n": 0, "start": None, "end": None, "ms": None}
    Base = tud.DataLoader

    class TimedLoader(Base):
        def __iter__(self):
            for batch in super().__iter__():
                if state["n"] == W:
                    torch.cuda.synchronize()
                    dist.barrier()
                    torch.cuda.synchronize()
                    state["start"] = torch.cuda.Event(enable_timing=True)
                    state["start"].record()
                if state["n"] == W + K:
                    state["end"] = torch.cuda.Event(enable_timing=True)
                    state["end"].record()
                    torch.cuda.synchronize()
                    dist.barrier()
                    state["ms"] = state["start"].elapsed_time(state["end"])
                    return  # enough: let the epoch (and dist_train) finish
                if state["ms"] is not None:
                    return
                state["n"] += 1
                yield batch

    ns = SimpleNamespace(gpus=world, epochs=epochs, backend="nccl", syncbn=args.syncbn, world_size=world,
                         init_method=f"tcp://127.0.0.1:{port}")
    tud.DataLoader = TimedLoader
    try:
        buf = io.StringIO()
        with redirect_stdout(buf):   # the reference prints every 10 steps; keep our stdout to one JSON line
            ref.dist_train(local_rank, ns)
    finally:
        tud.DataLoader = Base
    if state["ms"] is None:
        raise RuntimeError(f"reference loop ended after {state['n']} steps, before warmup+steps={W + K}")
    t = torch.tensor([state["ms"]], device=torch.device("cuda", local_rank))
    dist.all_reduce(t, dist.ReduceOp.MAX)
    ms = float(t.item())
    dist.destroy_process_group()
    return ms


def main():
    args = parse()
    if args.impl == "reference":
        out = run_reference(args)
    else:
        out = run_ours(args)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
