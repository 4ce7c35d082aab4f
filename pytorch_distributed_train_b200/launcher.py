"""One-process-per-GPU launcher.

Replaces the ``mp.spawn(dist_train, nprocs=args.gpus, args=(args,))`` call of the reference
(ref: ddp_example.py:111).  Contract (torch/multiprocessing/spawn.py:79-96,145-211): child *i*
runs ``fn(i, *args)`` in a fresh interpreter (``spawn`` start method, no inherited CUDA
context); the parent blocks; the first child failure terminates the siblings and is re-raised in
the parent with the child's traceback; a child killed by a signal is reported with the signal
name; children die with the parent.

Our own additions: a configurable kill grace period, ``PDT_LAUNCH_ID`` exported to children so
per-run resources (AF_UNIX names for VMM-handle passing) never collide between concurrent jobs,
and per-rank CPU affinity hints.
"""
from __future__ import annotations

import multiprocessing as mp
import multiprocessing.connection
import os
import pickle
import signal
import sys
import tempfile
import time
import traceback
import uuid
from typing import Callable, Dict, Optional, Sequence, Tuple


class ProcessException(Exception):
    def __init__(self, msg: str, error_index: int, pid: int):
        super().__init__(msg)
        self.msg = msg
        self.error_index = error_index
        self.pid = pid

    def __reduce__(self):
        return type(self), (self.msg, self.error_index, self.pid)


class ProcessRaisedException(ProcessException):
    """A child raised a Python exception; ``msg`` carries its formatted traceback."""


class ProcessExitedException(ProcessException):
    """A child exited abnormally (non-zero exit code or signal) without a Python traceback."""

    def __init__(self, msg: str, error_index: int, pid: int, exit_code: int, signal_name: Optional[str] = None):
        super().__init__(msg, error_index, pid)
        self.exit_code = exit_code
        self.signal_name = signal_name

    def __reduce__(self):
        return type(self), (self.msg, self.error_index, self.pid, self.exit_code, self.signal_name)


def _set_pdeathsig(sig: int) -> None:
    """Ask the kernel to signal us when the parent dies (Linux prctl PR_SET_PDEATHSIG)."""
    try:
        import ctypes

        libc = ctypes.CDLL(None, use_errno=True)
        libc.prctl(1, sig, 0, 0, 0)  # PR_SET_PDEATHSIG = 1
    except Exception:  # non-Linux: best effort
        pass


def _child_main(fn: Callable, index: int, args: Tuple, error_file: str, env: Dict[str, str]) -> None:
    _set_pdeathsig(signal.SIGINT)
    os.environ.update(env)
    try:
        fn(index, *args)
    except KeyboardInterrupt:
        pass  # parent died or user hit ^C: exit quietly
    except BaseException:  # noqa: BLE001 - the whole point is to ship *any* failure upstream
        with open(error_file, "wb") as f:
            pickle.dump(traceback.format_exc(), f)
        sys.exit(1)


class ProcessContext:
    def __init__(self, processes, error_files, grace_period: float = 30.0):
        self.processes = processes
        self.error_files = error_files
        self.grace_period = grace_period
        self.sentinels = {p.sentinel: i for i, p in enumerate(processes)}

    def pids(self):
        return [int(p.pid) for p in self.processes]

    def _kill_rest(self, except_index: int):
        for i, p in enumerate(self.processes):
            if i != except_index and p.is_alive():
                p.terminate()
        deadline = time.monotonic() + self.grace_period
        for i, p in enumerate(self.processes):
            if i == except_index:
                continue
            p.join(max(0.0, deadline - time.monotonic()))
        for i, p in enumerate(self.processes):
            if i != except_index and p.is_alive():
                p.kill()
                p.join()

    def join(self, timeout: Optional[float] = None) -> bool:
        """Wait for *one* more child to finish. Returns True when all are done.
        Raises on the first failure after terminating the siblings."""
        if not self.sentinels:
            return True
        ready = multiprocessing.connection.wait(list(self.sentinels.keys()), timeout=timeout)
        error_index = None
        for s in ready:
            idx = self.sentinels.pop(s)
            p = self.processes[idx]
            p.join()
            if p.exitcode != 0 and error_index is None:
                error_index = idx
        if error_index is None:
            return not self.sentinels
        self._kill_rest(error_index)
        failed = self.processes[error_index]
        err_path = self.error_files[error_index]
        if os.path.exists(err_path) and os.path.getsize(err_path) > 0:
            with open(err_path, "rb") as f:
                tb = pickle.load(f)
            msg = f"\n\n-- Process {error_index} terminated with the following error:\n{tb}"
            raise ProcessRaisedException(msg, error_index, failed.pid)
        code = failed.exitcode
        if code is not None and code < 0:
            try:
                name = signal.Signals(-code).name
            except ValueError:
                name = f"<unknown signal {-code}>"
            raise ProcessExitedException(f"process {error_index} terminated with signal {name}", error_index,
                                         failed.pid, code, name)
        raise ProcessExitedException(f"process {error_index} terminated with exit code {code}", error_index,
                                     failed.pid, code)

    def cleanup(self):
        for f in self.error_files:
            try:
                os.unlink(f)
            except OSError:
                pass


def start_processes(fn: Callable, args: Sequence = (), nprocs: int = 1, join: bool = True, daemon: bool = False,
                    start_method: str = "spawn", grace_period: float = 30.0, env: Optional[Dict[str, str]] = None):
    if nprocs < 1:
        raise ValueError("nprocs must be >= 1")
    ctx = mp.get_context(start_method)
    launch_id = os.environ.get("PDT_LAUNCH_ID") or uuid.uuid4().hex[:12]
    tmpdir = tempfile.mkdtemp(prefix="pdt_spawn_")
    processes, error_files = [], []
    for i in range(nprocs):
        err = os.path.join(tmpdir, f"rank{i}.err")
        child_env = {"PDT_LAUNCH_ID": launch_id, "PDT_LOCAL_RANK": str(i), "PDT_LOCAL_WORLD_SIZE": str(nprocs)}
        if env:
            child_env.update(env)
        p = ctx.Process(target=_child_main, args=(fn, i, tuple(args), err, child_env), daemon=daemon)
        p.start()  # sequential start, like the reference's substrate
        processes.append(p)
        error_files.append(err)
    context = ProcessContext(processes, error_files, grace_period)
    if not join:
        return context
    try:
        while not context.join():
            pass
    finally:
        context.cleanup()
        try:
            os.rmdir(tmpdir)
        except OSError:
            pass
    return None


def spawn(fn: Callable, args: Sequence = (), nprocs: int = 1, join: bool = True, daemon: bool = False,
          start_method: str = "spawn", **kw):
    """``spawn(fn, args=(...), nprocs=N)`` → child *i* runs ``fn(i, *args)``."""
    if start_method != "spawn":
        raise ValueError("spawn() only supports start_method='spawn' (CUDA contexts do not survive fork); "
                         "use start_processes() for other start methods")
    return start_processes(fn, args, nprocs, join, daemon, start_method="spawn", **kw)


# ---- script launcher: `python -m pytorch_distributed_train_b200.launcher --nproc-per-node N script.py args…` -------------
def _run_script(local_rank: int, script: str, script_args: Tuple[str, ...], as_module: bool) -> None:
    import runpy

    sys.argv = [script] + list(script_args)
    try:
        if as_module:
            runpy.run_module(script, run_name="__main__", alter_sys=True)
        else:
            runpy.run_path(script, run_name="__main__")
    except SystemExit as e:  # a script that ends with sys.exit(0) succeeded
        if e.code not in (None, 0):
            raise


def run(argv: Optional[Sequence[str]] = None) -> None:
    """Single-node equivalent of ``torchrun --standalone``: starts ``--nproc-per-node`` copies of a script with
    RANK / LOCAL_RANK / WORLD_SIZE / LOCAL_WORLD_SIZE / MASTER_ADDR / MASTER_PORT exported, so a script that calls
    ``init_process_group(init_method="env://")`` (ours or torch's) works unchanged; the same failure contract as
    ``spawn`` (first failing rank kills its siblings, its traceback is re-raised, exit status is non-zero)."""
    import argparse
    import socket

    p = argparse.ArgumentParser(prog="python -m pytorch_distributed_train_b200.launcher",
                                description="one process per GPU on this node, env:// rendezvous")
    p.add_argument("--nproc-per-node", "--nproc_per_node", type=int, default=1)
    p.add_argument("--master-addr", "--master_addr", default="127.0.0.1")
    p.add_argument("--master-port", "--master_port", type=int, default=0, help="0 = pick a free port")
    p.add_argument("-m", "--module", action="store_true", help="treat SCRIPT as a module name (python -m)")
    p.add_argument("script")
    p.add_argument("script_args", nargs=argparse.REMAINDER)
    a = p.parse_args(argv)
    port = a.master_port
    if port == 0:
        with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
            s.bind((a.master_addr, 0))
            port = s.getsockname()[1]
    n = a.nproc_per_node
    ctx = mp.get_context("spawn")
    tmpdir = tempfile.mkdtemp(prefix="pdt_run_")
    launch_id = uuid.uuid4().hex[:12]
    procs, errs = [], []
    for i in range(n):
        env = {"RANK": str(i), "LOCAL_RANK": str(i), "WORLD_SIZE": str(n), "LOCAL_WORLD_SIZE": str(n), "MASTER_ADDR": a.master_addr,
               "MASTER_PORT": str(port), "PDT_LAUNCH_ID": launch_id, "PDT_LOCAL_RANK": str(i), "PDT_LOCAL_WORLD_SIZE": str(n)}
        err = os.path.join(tmpdir, f"rank{i}.err")
        pr = ctx.Process(target=_child_main, args=(_run_script, i, (a.script, tuple(a.script_args), a.module), err, env))
        pr.start()
        procs.append(pr)
        errs.append(err)
    context = ProcessContext(procs, errs, grace_period=30.0)
    try:
        while not context.join():
            pass
    finally:
        context.cleanup()
        try:
            os.rmdir(tmpdir)
        except OSError:
            pass


if __name__ == "__main__":
    try:
        run(sys.argv[1:])
    except ProcessException as e:
        sys.stderr.write(str(e) + "\n")
        sys.exit(1)
