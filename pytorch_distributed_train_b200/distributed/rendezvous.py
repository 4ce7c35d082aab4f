"""Rendezvous: turn an ``init_method`` URL into a connected :class:`Store`.

Surface parity with what the reference reaches through
``dist.init_process_group(init_method='tcp://10.9.1.2:34567', world_size=, rank=)``
(ref: ddp_example.py:50,110; torch/distributed/rendezvous.py:210-239 tcp, :242-287 env, :126-155
file).  Rank 0 hosts the store server, every other rank dials it with retry, so start order
does not matter.

``env://`` additionally understands ``torchrun``: the elastic agent already owns
``MASTER_PORT`` for its own store, so our store binds ``PDT_STORE_PORT`` if set, else
``MASTER_PORT + 1`` with a same-host handshake file as a fallback when that port is taken.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple
from urllib.parse import parse_qs, urlparse

from .. import _C

DEFAULT_PORT = 29500


class RendezvousError(RuntimeError):
    pass


def _query_int(q, key) -> Optional[int]:
    if key in q:
        try:
            return int(q[key][0])
        except ValueError as e:
            raise RendezvousError(f"init_method query parameter {key}={q[key][0]!r} is not an integer") from e
    return None


def _tcp(url, rank, world_size, timeout):
    if not url.hostname or url.port is None:
        raise RendezvousError(f"tcp:// init_method needs host and port, got {url.geturl()!r}")
    q = parse_qs(url.query)
    rank = _query_int(q, "rank") if rank is None or rank < 0 else rank
    world_size = _query_int(q, "world_size") if world_size is None or world_size < 0 else world_size
    if rank is None or world_size is None:
        raise RendezvousError("tcp:// rendezvous needs rank and world_size (arguments or ?rank=&world_size=)")
    store = _C.TCPStore(url.hostname, url.port, world_size, rank == 0, timeout, True)
    return store, rank, world_size


def _file(url, rank, world_size, timeout):
    path = url.path
    if url.netloc and url.netloc not in ("localhost",):
        path = "/" + url.netloc + path  # tolerate file://tmp/x meaning /tmp/x
    if not path:
        raise RendezvousError("file:// init_method needs a path")
    q = parse_qs(url.query)
    rank = _query_int(q, "rank") if rank is None or rank < 0 else rank
    world_size = _query_int(q, "world_size") if world_size is None or world_size < 0 else world_size
    if rank is None or world_size is None:
        raise RendezvousError("file:// rendezvous needs rank and world_size")
    store = _C.FileStore(path, world_size)
    store.set_timeout(timeout)
    return store, rank, world_size


def _env_int(name, override=None) -> int:
    if override is not None and override >= 0:
        return override
    v = os.environ.get(name)
    if v is None:
        raise RendezvousError(f"env:// rendezvous: environment variable {name} is not set")
    return int(v)


def _env(url, rank, world_size, timeout):
    q = parse_qs(url.query) if url is not None else {}
    rank = _env_int("RANK", _query_int(q, "rank") if rank is None or rank < 0 else rank)
    world_size = _env_int("WORLD_SIZE", _query_int(q, "world_size") if world_size is None or world_size < 0 else world_size)
    addr = os.environ.get("MASTER_ADDR")
    port = os.environ.get("MASTER_PORT")
    if addr is None or port is None:
        raise RendezvousError("env:// rendezvous: MASTER_ADDR and MASTER_PORT must be set")
    port = int(port)
    under_torchrun = "TORCHELASTIC_RUN_ID" in os.environ or "TORCHELASTIC_RESTART_COUNT" in os.environ
    if "PDT_STORE_PORT" in os.environ:
        port = int(os.environ["PDT_STORE_PORT"])
    elif under_torchrun:
        return _torchrun_store(addr, port, rank, world_size, timeout)
    store = _C.TCPStore(addr, port, world_size, rank == 0, timeout, True)
    return store, rank, world_size


def _torchrun_store(addr, master_port, rank, world_size, timeout):
    """The elastic agent holds MASTER_PORT; put our store on MASTER_PORT+1.. and publish the
    port we really got in a same-host handshake file so peers never guess."""
    import time

    run_id = os.environ.get("TORCHELASTIC_RUN_ID", "none")
    restart = os.environ.get("TORCHELASTIC_RESTART_COUNT", "0")
    hs = f"/tmp/pdt_rdzv_{addr}_{master_port}_{run_id}_{restart}"
    if rank == 0:
        last = None
        for cand in [master_port + 1 + i for i in range(32)] + [0]:
            try:
                store = _C.TCPStore(addr, cand, world_size, True, timeout, False)
                break
            except RuntimeError as e:  # port taken
                last = e
        else:
            raise RendezvousError(f"could not bind a store port near {master_port}: {last}")
        tmp = f"{hs}.{os.getpid()}"
        with open(tmp, "w") as f:
            f.write(f"{store.port} {os.getpid()}")
        os.replace(tmp, hs)
        # wait for the workers ourselves (we created the server with wait_for_workers=False so the
        # handshake file could be published first)
        deadline = time.time() + timeout
        while store.add("__pdt_store_init__/workers", 0) < world_size:
            if time.time() > deadline:
                raise RendezvousError(f"rank 0 timed out waiting for {world_size} workers to join the store")
            time.sleep(0.002)
        return store, rank, world_size
    deadline = time.time() + timeout
    last = None
    while time.time() < deadline:
        try:
            with open(hs) as f:
                port_s, pid_s = f.read().split()
            # a stale file from an earlier run points at a dead pid: skip it
            if os.path.exists(f"/proc/{pid_s}"):
                store = _C.TCPStore(addr, int(port_s), world_size, False, min(5.0, timeout), False)
                store.set_timeout(timeout)
                return store, rank, world_size
        except (OSError, ValueError, TimeoutError, RuntimeError) as e:
            last = e
        time.sleep(0.01)
    raise RendezvousError(f"rank {rank}: could not find rank 0's store via {hs}: {last}")


def rendezvous(init_method: Optional[str], rank: int = -1, world_size: int = -1,
               timeout: float = 300.0) -> Tuple["_C.Store", int, int]:
    """Returns ``(store, rank, world_size)`` for ``tcp://host:port``, ``env://`` or ``file:///path``."""
    if init_method is None:
        init_method = "env://"
    url = urlparse(init_method)
    if url.scheme == "tcp":
        return _tcp(url, rank, world_size, timeout)
    if url.scheme == "env":
        return _env(url, rank, world_size, timeout)
    if url.scheme == "file":
        return _file(url, rank, world_size, timeout)
    raise RendezvousError(f"unsupported init_method scheme {url.scheme!r} (supported: tcp://, env://, file://)")
