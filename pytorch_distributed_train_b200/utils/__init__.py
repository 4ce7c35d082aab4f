"""Auxiliary subsystems: logging, device timers, NVTX ranges, clock sampling, watchdog, checkpoints."""
from .checkpoint import load_checkpoint, save_checkpoint
from .clocks import ClockSampler
from .logging import get_logger, rank_zero_print
from .nvtx import nvtx_range
from .timers import DeviceTimer, l2_flush, max_over_ranks
from .watchdog import Watchdog

__all__ = ["ClockSampler", "get_logger", "rank_zero_print", "nvtx_range", "DeviceTimer", "l2_flush",
           "max_over_ranks", "Watchdog", "save_checkpoint", "load_checkpoint"]
