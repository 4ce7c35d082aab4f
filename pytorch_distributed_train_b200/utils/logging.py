"""Rank-aware logging. The reference prints with bare ``print`` on rank 0
(ref: ddp_example.py:49,55-56,93-97); the training CLI keeps those exact formats and uses
this module for everything else."""
from __future__ import annotations

import logging
import os
import sys

_FMT = "[%(asctime)s][pdt][rank%(rank)s] %(levelname)s %(message)s"


class _RankFilter(logging.Filter):
    def filter(self, record):
        from .. import distributed as dist

        record.rank = dist.get_rank() if dist.is_initialized() else os.environ.get("RANK", "-")
        return True


def get_logger(name: str = "pdt") -> logging.Logger:
    log = logging.getLogger(name)
    if not log.handlers:
        h = logging.StreamHandler(sys.stderr)
        h.setFormatter(logging.Formatter(_FMT, "%H:%M:%S"))
        h.addFilter(_RankFilter())
        log.addHandler(h)
        log.setLevel(os.environ.get("PDT_LOG_LEVEL", "WARNING").upper())
        log.propagate = False
    return log


def rank_zero_print(*args, **kw):
    from .. import distributed as dist

    if not dist.is_initialized() or dist.get_rank() == 0:
        print(*args, **kw)
        sys.stdout.flush()
