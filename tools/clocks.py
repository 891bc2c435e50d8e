"""tools/clocks.py -- SM clock / throttle-reason sampling DURING a timed region (bench infra).

Samples NVML every few milliseconds from a background thread; only samples whose timestamp lies
inside [begin(), end()] are reported.  Falls back to an `nvidia-smi -lms` subprocess when the
NVML bindings are missing (the recipe of B200_PROFILING.md)."""
from __future__ import annotations

import statistics
import subprocess
import threading
import time


class ClockSampler:
    REASONS = {  # nvmlClocksEventReasons bit masks
        "hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
        "sw_power_cap": 0x4, "hw_power_brake": 0x80, "sync_boost": 0x10,
    }

    def __init__(self, gpu_index: int, period_s: float = 0.004):
        self.idx = gpu_index
        self.period = period_s
        self.samples = []       # (t, sm_mhz, power_w, reasons_mask)
        self.t0 = self.t1 = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        self.sm_max = None
        self._smi = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None
            try:
                self._smi = subprocess.Popen(
                    ["nvidia-smi", f"--id={self.idx}",
                     "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                     "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                     "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "20"],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self._thread = threading.Thread(target=self._pump, daemon=True)
                self._thread.start()
            except Exception:
                self._smi = None

    def _loop(self):
        n = self._nvml
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
                pw = n.nvmlDeviceGetPowerUsage(self._h) / 1000.0
                try:
                    rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                self.samples.append((time.perf_counter(), sm, pw, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def _pump(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self._smi.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                mask = 0
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        mask |= self.REASONS[nm]
                self.sm_max = float(f[1])
                self.samples.append((time.perf_counter(), float(f[0]), float(f[2]), mask))
            except Exception:
                continue

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self._stop.set()
        if self._smi is not None:
            self._smi.terminate()
        if self._thread is not None:
            self._thread.join(timeout=2)
        if self._nvml is None and self._smi is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        inside = [s for s in self.samples if self.t0 is not None and self.t0 <= s[0] <= (self.t1 or 1e30)]
        note = None
        if not inside:          # region shorter than one sampling period: use the nearest samples
            inside = self.samples[-3:]
            note = "timed region shorter than the sampling period; nearest samples used"
        mask = 0
        for s in inside:
            mask |= s[3]
        out = {"sm_mhz": statistics.median(s[1] for s in inside) if inside else None,
               "sm_max_mhz": self.sm_max,
               "power_w_max": max((s[2] for s in inside), default=None),
               "samples": len(inside),
               "reasons": sorted(k for k, b in self.REASONS.items() if mask & b),
               "source": "nvml" if self._nvml is not None else "nvidia-smi"}
        if note:
            out["note"] = note
        return out
