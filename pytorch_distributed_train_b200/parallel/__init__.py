"""Parallelism strategies: DistributedDataParallel (the product), SyncBatchNorm, DataParallel, comm hooks."""
from . import comm_hooks
from .data_parallel import DataParallel
from .ddp import DistributedDataParallel, broadcast_coalesced
from .sync_batchnorm import SyncBatchNorm

__all__ = ["DistributedDataParallel", "SyncBatchNorm", "DataParallel", "broadcast_coalesced", "comm_hooks"]
