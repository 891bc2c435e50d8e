"""Test helper (run by hand on a GPU box: python tests/stress_tree.py 150): stress test: repeat small tree traversals with poisoned device memory and report any
mismatch against the oracle (site list, which categories, tuning)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # tests/ -> repo root
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, oracle, torch
from oracle import tree_oracle
from test_tree import tree_inputs
pkg = bench.load_pkg()
co = oracle.COracle()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cases = []
for shape, n_tips, n in (("random", 257, 517), ("random", 33, 1000), ("balanced", 64, 3001)):
    left, right = (pkg.random_tree(n_tips, seed=n) if shape == "random" else pkg.balanced_tree(n_tips))
    tips, ev, pl, pr, wgt = tree_inputs(n_tips, n, seed=n_tips)
    cases.append((shape, n_tips, n, left, right, tips, ev, pl, pr, wgt, tree_oracle.traverse(co, left, right, tips, ev, pl, pr, wgt)))
fails = 0
t0 = time.time()
for it in range(iters):
    poison = torch.full((64 << 20,), float("nan"), device="cuda")   # 256 MiB of NaN into the allocator's pool
    del poison
    torch.cuda.empty_cache()
    for shape, n_tips, n, left, right, tips, ev, pl, pr, wgt, (o_root, o_cnt, o_total) in cases:
        for u, chunk in ((1, 1000), (2, 1000), (1, 1), (0, 0)):
            with pkg.Tree(left, right, n) as t:
                t.set_tuning(u, chunk)
                for i in range(n_tips):
                    t.write_tip(i, tips[i])
                t.write_matrices(ev, pl, pr)
                t.write_wgt(wgt)
                for rep in range(2):
                    t.run_async()
                    root, cnt = t.read_root()
                    neq = root.view(np.uint32) != o_root.view(np.uint32)
                    if neq.any() or (cnt != o_cnt).any() or t.total_scalings() != o_total:
                        fails += 1
                        bad = np.nonzero(neq.any(axis=1))[0]
                        print("MISMATCH it", it, shape, n_tips, n, "u", u, "chunk", chunk, "rep", rep, "sites", bad[:10].tolist(),
                              "elems of first", np.nonzero(neq[bad[0]])[0].tolist() if len(bad) else None,
                              "cnt_bad", int((cnt != o_cnt).sum()), flush=True)
print("iterations", iters, "failures", fails, "seconds", round(time.time() - t0, 1))
