#!/bin/bash
# f1: ensi_multi + staticcorr tests, C++ API test; then timing of the new kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_cxx_api.py -x -q -m gpu -k "ensi_multi or cxx" 2>&1 | tail -30 > gpurun_out/r2_pytest_f1.log
cat gpurun_out/r2_pytest_f1.log
