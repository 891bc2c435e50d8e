#!/bin/bash
# three-way overlap of the streamed host path (no nsys in the image: CUDA events per chunk)
set -u
mkdir -p gpurun_out
P=amd-versal-phylogenetic-likelihood-function_b200
CFG=plf_128x9DNAwindow8192Comb_memDNAwindowComb
timeout 300 $P/host_stream.exe $CFG 0 16777216 2 > gpurun_out/c32_host_stream.txt 2>&1; echo "plain rc=$?"; tail -12 gpurun_out/c32_host_stream.txt
PLF_STREAM_TRACE=gpurun_out/c32_stream_trace.txt NO_CORRECTNESS_CHECK=1 timeout 300 $P/host_stream.exe $CFG 0 16777216 2 > gpurun_out/c32_host_stream_traced.txt 2>&1; echo "traced rc=$?"
python tools/stream_timeline.py gpurun_out/c32_stream_trace.txt > gpurun_out/c32_stream_timeline.txt 2>&1; cat gpurun_out/c32_stream_timeline.txt
