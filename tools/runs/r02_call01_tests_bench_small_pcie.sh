#!/bin/bash
# round-2 GPU call 1 (one B200): tests, the bench line with side workloads, small-launch study, PCIe probe.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
nproc > gpurun_out/c1_host.txt; free -g >> gpurun_out/c1_host.txt; lscpu | head -25 >> gpurun_out/c1_host.txt
numactl -H >> gpurun_out/c1_host.txt 2>&1; cat /sys/kernel/mm/transparent_hugepage/enabled >> gpurun_out/c1_host.txt 2>&1
grep -i huge /proc/meminfo >> gpurun_out/c1_host.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench_n1.json 2> gpurun_out/c1_bench_n1.err; echo "bench rc=$?"
timeout 300 python tools/small_launch.py --out gpurun_out/c1_small.json > gpurun_out/c1_small.log 2>&1; echo "small rc=$?"
for mode in h2d d2h duplex; do for alloc in pinned thp huge; do
  timeout 120 build/pcie_probe --gpus 0 --mode $mode --alloc $alloc --mb 1024 --reps 6 >> gpurun_out/c1_pcie.jsonl 2>> gpurun_out/c1_pcie.err
done; done
timeout 120 build/pcie_probe --gpus 0 --mode duplex --alloc pinned --mb 1024 --reps 6 --streams 4 >> gpurun_out/c1_pcie.jsonl 2>> gpurun_out/c1_pcie.err
timeout 120 build/pcie_probe --gpus 0 --mode duplex --alloc pinned --mb 64 --reps 64 >> gpurun_out/c1_pcie.jsonl 2>> gpurun_out/c1_pcie.err
for rel in 0 1; do
  PLF_SAFE_RELEASE=$rel timeout 300 python tools/sweep.py --sites 67108864 --reps 20 --only 1432:512:0:0,1432:512:0:1,1422:512:0:0 --out gpurun_out/c1_sweep64_rel$rel.json > gpurun_out/c1_sweep64_rel$rel.log 2>&1
  PLF_SAFE_RELEASE=$rel timeout 300 python tools/sweep.py --sites 8388608 --reps 40 --only 1432:512:0:0,1432:512:0:1,2334:256:0:0,2632:256:0:0 --out gpurun_out/c1_sweep8_rel$rel.json > gpurun_out/c1_sweep8_rel$rel.log 2>&1
done
cfg=plf_128x9DNAwindow8192Comb_memDNAwindowComb
( cd amd-versal-phylogenetic-likelihood-function_b200 && timeout 300 ./host_mem.exe $cfg 0 1048576 20 9 > ../gpurun_out/c1_host_mem_1Mi_9inst.txt 2>&1; timeout 300 ./host_mem.exe $cfg 0 1048576 20 1 > ../gpurun_out/c1_host_mem_1Mi_1inst.txt 2>&1 )
echo done
