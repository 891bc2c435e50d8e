"""Developer helper: a few launches of one 20-state kernel shape for an ncu capture.
    python tools/protein_prof.py <variant> <threads> <math 0|1> [sites]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import bench
import protein_bench
pkg = bench.load_pkg()
v, t, m = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 21
protein_bench.measure(pkg, torch, n, 1, shapes=[(v, t)], maths=(m,))
