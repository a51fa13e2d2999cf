#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "edge_cases or cholesky" 2>&1 | tail -25 | cut -c1-300
