"""Developer repro for tree traversal mismatches: prints which sites differ for several tunings."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, oracle
from oracle import tree_oracle
from test_tree import tree_inputs
pkg = bench.load_pkg()
co = oracle.COracle()
for shape, n_tips, n in (("random", 257, 517), ("balanced", 257, 517), ("random", 257, 512), ("random", 257, 640), ("random", 100, 517)):
    left, right = (pkg.random_tree(n_tips, seed=n) if shape == "random" else pkg.balanced_tree(n_tips))
    tips, ev, pl, pr, wgt = tree_inputs(n_tips, n, seed=n_tips)
    o_root, o_cnt, o_total = tree_oracle.traverse(co, left, right, tips, ev, pl, pr, wgt)
    for u, chunk in ((1, 1000), (1, 1000), (2, 1000), (1, 1), (1, 7), (1, 64), (2, 64), (0, 0)):
        with pkg.Tree(left, right, n) as t:
            t.set_tuning(u, chunk)
            for i in range(n_tips):
                t.write_tip(i, tips[i])
            t.write_matrices(ev, pl, pr)
            t.write_wgt(wgt)
            t.run_async()
            root, cnt = t.read_root()
            bad = np.nonzero((root.view(np.uint32) != o_root.view(np.uint32)).any(axis=1))[0]
            print(shape, n_tips, n, "u", u, "chunk", chunk, "levels", t.info()["levels"], "bad sites", len(bad), bad[:12].tolist(),
                  "cnt_bad", int((cnt != o_cnt).sum()), "total", t.total_scalings() == o_total, flush=True)
