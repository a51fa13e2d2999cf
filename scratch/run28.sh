#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python profiles/oi_slices.py 1 2 4 8 2>&1 | tail -4 | tee gpurun_out/r2_oi_slices_v5.log
timeout 900 python -m pytest tests -x -q -m gpu -k "oi or optimal" 2>&1 | tail -3
