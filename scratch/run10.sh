#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu.py -x -q -m gpu -k "oi or optimal" 2>&1 | tail -5
for fl in 0 1; do echo "first level $fl"; GPP_OI_FIRST_LEVEL=$fl timeout 600 python profiles/oi_slices.py 1 2 4 8 2>&1 | tail -5; done | tee gpurun_out/r2_oi_slices_v3.log
echo default; timeout 600 python profiles/oi_slices.py 1 4 8 2>&1 | tail -5 | tee -a gpurun_out/r2_oi_slices_v3.log
