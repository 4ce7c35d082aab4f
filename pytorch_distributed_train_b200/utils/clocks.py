"""Sample SM clocks / throttle reasons with ``nvidia-smi`` *during* a timed region
(B200_PROFILING.md "clocks line").  A background thread runs one long-lived
``nvidia-smi -lms`` process; ``summary()`` gives the JSON block bench.py prints."""
from __future__ import annotations

import shutil
import statistics
import subprocess
import threading
from typing import List, Optional

_Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
      "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
      "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    def __init__(self, gpu_index: Optional[int] = None, period_ms: int = 100):
        self.gpu_index = gpu_index
        self.period_ms = period_ms
        self.rows: List[List[str]] = []
        self.proc = None
        self.thread = None

    def start(self):
        exe = shutil.which("nvidia-smi")
        if exe is None:
            return self
        cmd = [exe, f"--query-gpu={_Q}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms)]
        if self.gpu_index is not None:
            cmd += ["-i", str(self.gpu_index)]
        try:
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return self

        def pump():
            for line in self.proc.stdout:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 9:
                    self.rows.append(parts)

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()
        return self

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            if self.thread is not None:
                self.thread.join(1)
        return self

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()
        return False

    def summary(self) -> dict:
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
            except ValueError:
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "reasons": sorted(reasons), "samples": len(sm)}
