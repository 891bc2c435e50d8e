#!/bin/bash
# time per launch against launch size (1 ... 64 Mi sites), four kernel shapes
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_protein_tc.py -m gpu -q > gpurun_out/c34_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c34_pytest.log
timeout 600 python tools/size_curve.py --shapes 0:0,1432:512,1332:512,1334:256 --out gpurun_out/c34_size_curve.json > gpurun_out/c34_size_curve.log 2>&1; echo "rc=$?"
cut -c1-700 gpurun_out/c34_size_curve.log
