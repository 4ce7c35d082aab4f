"""Helpers to run a function on N ranks through our own launcher + rendezvous."""
import os
import pickle
import socket
import tempfile

import pytorch_distributed_train_b200 as pdt


def free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, fn, world_size, init_method, backend, outdir, args):
    import torch

    torch.set_num_threads(1)
    if backend not in ("gloo", "cpu"):
        torch.cuda.set_device(rank)
    pdt.init_process_group(backend=backend, init_method=init_method, world_size=world_size, rank=rank, timeout=60.0)
    try:
        res = fn(rank, world_size, *args)
    finally:
        pdt.destroy_process_group()
    with open(os.path.join(outdir, f"r{rank}.pkl"), "wb") as f:
        pickle.dump(res, f)


def run_ranks(fn, world_size, *args, backend="gloo", grace_period=5.0):
    """Runs fn(rank, world_size, *args) on every rank; returns the list of return values."""
    init = f"tcp://127.0.0.1:{free_port()}"
    with tempfile.TemporaryDirectory() as d:
        pdt.spawn(_entry, args=(fn, world_size, init, backend, d, args), nprocs=world_size, grace_period=grace_period)
        out = []
        for r in range(world_size):
            with open(os.path.join(d, f"r{r}.pkl"), "rb") as f:
                out.append(pickle.load(f))
        return out
