#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for E in 20 12 16 24 25; do python profiles/ensi_device_time.py 1000 1 $E 2>&1 | tail -1; done
timeout 900 python -m pytest tests -x -q -m gpu -k "ensi or utem or edge_cases" 2>&1 | tail -3
