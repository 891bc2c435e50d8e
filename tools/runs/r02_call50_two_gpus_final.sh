#!/bin/bash
# final two-GPU check of the round: multi-GPU tests (DNA and 20 states through plf_multi_*, NCCL totals), host_mem on two
# GPUs with both state counts, the N = 2 bench line and its reference arm
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi.py tests/test_states_api.py -m gpu -q > gpurun_out/c50_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c50_pytest.log
tail -3 gpurun_out/c50_pytest.log
( cd amd-versal-phylogenetic-likelihood-function_b200 && NCCL_DEBUG=WARN timeout 300 ./host_mem.exe plf_128x9DNAwindow8192Comb_memDNAwindowComb 0,1 4194304 3 18 > ../gpurun_out/c50_host_mem_2gpu_dna.txt 2>&1; echo "host_mem DNA rc=$?"
  NCCL_DEBUG=WARN timeout 300 ./host_mem.exe plf_128x9AAwindow8192Comb_memAAwindowComb 0,1 1000000 3 18 > ../gpurun_out/c50_host_mem_2gpu_aa.txt 2>&1; echo "host_mem AA rc=$?" )
grep -h "Test result\|scalerIncrement" gpurun_out/c50_host_mem_2gpu_dna.txt gpurun_out/c50_host_mem_2gpu_aa.txt | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c50_bench_n2.json 2> gpurun_out/c50_bench_n2.err; echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/c50_bench_ref_n2.json 2> gpurun_out/c50_bench_ref_n2.err; echo "reference arm rc=$?"
