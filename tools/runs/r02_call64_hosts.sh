#!/bin/bash
# host program tests after host_stream.exe gained the STATES knob; one 2 M-site AA run
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hosts.py -m gpu -q > gpurun_out/c64_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c64_pytest.log
timeout 300 amd-versal-phylogenetic-likelihood-function_b200/host_stream.exe plf_128x9AAwindow8192Comb_memAAwindowComb 0 2000000 2 > gpurun_out/c64_host_stream_aa.txt 2>&1; echo "host_stream AA rc=$?"; tail -9 gpurun_out/c64_host_stream_aa.txt
