"""The analysis tools under tools/ keep working on the samples committed under profiles/ (CPU only): the tensor-core
kernel's wait counters, the streamed path's event timeline, the ncu target list."""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    r = subprocess.run([sys.executable, *args], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_tc_trace_reads_the_committed_counters():
    out = run("tools/tc_trace.py", "profiles/r02_protein_tc_wait_counters.txt")
    assert "1184 worker warps, 16384 tiles" in out                   # 148 CTAs x 8 worker warps, 2 Mi sites / 128
    for key in ("x1 box (TMA)", "branch MMAs (mma_ab)", "EV MMAs (mma_x)", "segment convert x1, x2", "hand-off: last worker arrive"):
        assert key in out, key
    step = float(out.split("cycles per step: mean ")[1].split(",")[0])
    assert 1500 < step < 4000


def test_stream_timeline_reads_the_committed_trace():
    out = run("tools/stream_timeline.py", "profiles/r02_stream_trace_16Mi.txt")
    assert "chunks 16, sites 16777216" in out
    h2d = float(out.split("H2D copies")[1].split("=")[1].split("%")[0])
    d2h = float(out.split("D2H copies")[1].split("=")[1].split("%")[0])
    assert h2d > 90.0 and 40.0 < d2h < 65.0                        # the link's two directions: 128 : 65 bytes per site
    assert "chunk  0 slot 0 |H" in out


def test_tools_compile():
    for name in sorted(os.listdir(os.path.join(ROOT, "tools"))):
        if name.endswith(".py"):
            run("-m", "py_compile", os.path.join("tools", name))
