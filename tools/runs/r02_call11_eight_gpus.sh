#!/bin/bash
# round-2 GPU call on EIGHT B200s: host-link ceiling per N (bare pinned copies), the bench line at N=8 and N=4 with its
# side workloads, the reference arm, the native multi-GPU host with NCCL.
set -u
mkdir -p gpurun_out
P=build/pcie_probe
nvidia-smi topo -m > gpurun_out/c11_topo.txt 2>&1
# each GPU alone (is every link the same?)
for g in 0 1 2 3 4 5 6 7; do timeout 60 $P --gpus $g --mode duplex --mb 512 --reps 6 >> gpurun_out/c11_pcie.jsonl 2>> gpurun_out/c11_pcie.err; done
# 2, 4, 8 GPUs at once: both directions, each direction alone, transparent huge pages
for set in 0,1 0,4 0,1,2,3 4,5,6,7 0,2,4,6 0,1,2,3,4,5,6,7; do
  for mode in duplex h2d d2h; do timeout 90 $P --gpus $set --mode $mode --mb 512 --reps 6 >> gpurun_out/c11_pcie.jsonl 2>> gpurun_out/c11_pcie.err; done
done
timeout 90 $P --gpus 0,1,2,3,4,5,6,7 --mode duplex --alloc thp --mb 512 --reps 6 >> gpurun_out/c11_pcie.jsonl 2>> gpurun_out/c11_pcie.err
timeout 90 $P --gpus 0,1,2,3,4,5,6,7 --mode duplex --mb 64 --reps 48 >> gpurun_out/c11_pcie.jsonl 2>> gpurun_out/c11_pcie.err
timeout 90 $P --gpus 0,1,2,3,4,5,6,7 --mode duplex --mb 512 --reps 6 --streams 2 >> gpurun_out/c11_pcie.jsonl 2>> gpurun_out/c11_pcie.err
echo "probe done"
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/c11_bench_n$N.json 2> gpurun_out/c11_bench_n$N.err; echo "bench N=$N rc=$?"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 8 --steps 5 --warmup 2 > gpurun_out/c11_bench_ref_n8.json 2> gpurun_out/c11_bench_ref_n8.err; echo "ref rc=$?"
cfg=plf_128x9DNAwindow8192Comb_memDNAwindowComb
( cd amd-versal-phylogenetic-likelihood-function_b200 && NCCL_DEBUG=WARN timeout 300 ./host_mem.exe $cfg 0,1,2,3,4,5,6,7 16777216 3 72 > ../gpurun_out/c11_host_mem_8gpu.txt 2>&1; echo "host_mem rc=$?" )
timeout 300 python -m pytest tests/test_multi.py -m gpu -q > gpurun_out/c11_pytest_multi.log 2>&1; tail -2 gpurun_out/c11_pytest_multi.log
echo done
