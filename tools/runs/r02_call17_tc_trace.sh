#!/bin/bash
# the first clock64 stamp trace of the tcgen05 kernel (one traced thread)
set -u
mkdir -p gpurun_out
PLF_TC_TRACE=gpurun_out/c17_tc_trace.txt timeout 120 python - <<'P'
import sys, os
sys.path.insert(0, os.getcwd())
import bench, torch, numpy as np
pkg = bench.load_pkg()
dev = torch.device("cuda", 0); n = 2 << 20
rng = np.random.RandomState(7)
ev, left, right = (rng.random_sample(k).astype(np.float32) for k in (400, 1600, 1600))
x1 = torch.empty(n * 80, device=dev); x2 = torch.empty(n * 80, device=dev); x3 = torch.empty(n * 80, device=dev)
sc = torch.empty(n, dtype=torch.uint8, device=dev); dsum = torch.zeros(1, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
pkg.generate_states_device(20, x1.data_ptr(), x2.data_ptr(), 0, n, 42, st)
opts = pkg.make_opts(1, 9, 0, 0, 0, pkg.LAUNCH_DEP_RELEASE)
for _ in range(2):
    pkg.newview_states_device(20, x1.data_ptr(), x2.data_ptr(), x3.data_ptr(), sc.data_ptr(), ev, left, right, None, n, dsum.data_ptr(), opts, st)
torch.cuda.synchronize()
print("ok")
P
wc -l gpurun_out/c17_tc_trace.txt
