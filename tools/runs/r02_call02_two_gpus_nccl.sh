#!/bin/bash
# round-2 GPU call 2 (two B200s): native multi-GPU path with NCCL, stress tests, 2-GPU PCIe probe and bench line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi.py tests/test_stress.py tests/test_felsenstein.py tests/test_evaluate.py -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -5 gpurun_out/c2_pytest.log
cfg=plf_128x9DNAwindow8192Comb_memDNAwindowComb
( cd amd-versal-phylogenetic-likelihood-function_b200 && NCCL_DEBUG=WARN timeout 300 ./host_mem.exe $cfg 0,1 4194304 5 18 > ../gpurun_out/c2_host_mem_2gpu.txt 2>&1; echo "host_mem rc=$?" )
for g in 0 1 0,1; do for alloc in pinned thp; do
  timeout 120 build/pcie_probe --gpus $g --mode duplex --alloc $alloc --mb 1024 --reps 6 >> gpurun_out/c2_pcie.jsonl 2>> gpurun_out/c2_pcie.err
done; done
timeout 120 build/pcie_probe --gpus 0,1 --mode h2d --alloc pinned --mb 1024 --reps 6 >> gpurun_out/c2_pcie.jsonl 2>> gpurun_out/c2_pcie.err
timeout 120 build/pcie_probe --gpus 0,1 --mode d2h --alloc pinned --mb 1024 --reps 6 >> gpurun_out/c2_pcie.jsonl 2>> gpurun_out/c2_pcie.err
timeout 120 build/pcie_probe --gpus 0,1 --mode duplex --alloc pinned --mb 1024 --reps 6 --stagger-us 20000 >> gpurun_out/c2_pcie.jsonl 2>> gpurun_out/c2_pcie.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c2_bench_n2.json 2> gpurun_out/c2_bench_n2.err; echo "bench rc=$?"
echo done
